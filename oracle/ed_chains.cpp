// TEST INFRASTRUCTURE ONLY.  Runs the reference's closed-source Edge Drawing library on raw 8-bit images exactly as
// LineDetector::DetectEdgeMap does (LineDetector.cc:855: DetectEdgesByED(srcImg, width, height, SOBEL_OPERATOR, 36, 8, 1.0))
// and dumps the edge chains (EdgeMap::segments, the input of LineFitting :884-900 and of the mEdgeIndex mask :857-866).
// Linked against /root/reference/Thirdparty/EDTest/EDLib.a where it lies (oracle/Makefile target `ed`); used by
// oracle/make_ed_golden.py to write tests/golden/ed_chains_small.npz.  Nothing of this travels to the GPU box.
//   usage: ed_chains W H N in.raw out.bin      in.raw = N images of W*H bytes
//   out.bin (int32): N, then per image: noSegments, per segment: noPixels, (r, c) * noPixels
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "EDLib.h"

int main(int argc, char** argv)
{
    if (argc != 6) { fprintf(stderr, "usage: %s W H N in.raw out.bin\n", argv[0]); return 2; }
    const int W = atoi(argv[1]), H = atoi(argv[2]), N = atoi(argv[3]);
    FILE* fi = fopen(argv[4], "rb");
    FILE* fo = fopen(argv[5], "wb");
    if (!fi || !fo) { perror("open"); return 1; }
    std::vector<unsigned char> im((size_t)W * H);
    std::vector<int> out;
    out.push_back(N);
    for (int i = 0; i < N; ++i) {
        if (fread(im.data(), 1, im.size(), fi) != im.size()) { fprintf(stderr, "short read\n"); return 1; }
        EdgeMap* map = DetectEdgesByED(im.data(), W, H, SOBEL_OPERATOR, 36, 8, 1.0);
        out.push_back(map->noSegments);
        for (int s = 0; s < map->noSegments; ++s) {
            out.push_back(map->segments[s].noPixels);
            for (int j = 0; j < map->segments[s].noPixels; ++j) {
                out.push_back(map->segments[s].pixels[j].r);
                out.push_back(map->segments[s].pixels[j].c);
            }
        }
        // (the EdgeMap is leaked on purpose: its destructor uses scalar delete on arrays)
    }
    fwrite(out.data(), sizeof(int), out.size(), fo);
    fclose(fo);
    fclose(fi);
    return 0;
}
