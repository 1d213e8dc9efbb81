// test_edge_drawing.cpp - runs the open Edge Drawing implementation (eao-slam_b200/host/edge_drawing.h) on raw 8-bit images
// and dumps the chains in the format of oracle/ed_chains.cpp (the same dump of the reference's closed-source EDLib.a), so that
// tests/test_edge_drawing.py can compare the two byte for byte.
//   usage: test_edge_drawing W H N in.raw out.bin [edge_index.bin [planes.bin]]      ("-" skips edge_index.bin)
//   planes.bin: per image the stage-1 planes of EdPlanesHost, G (int16 W*H) then F (uint8 W*H)
//   prints the host time of the detector per image (both stages, one thread) on stderr
//   out.bin (int32): N, then per image: noSegments, per segment: noPixels, (r, c) * noPixels
#include <stdio.h>
#include <time.h>
#include <stdlib.h>

#include <vector>

#include "../../eao-slam_b200/host/edge_drawing.h"

int main(int argc, char** argv)
{
    if (argc < 6) { fprintf(stderr, "usage: %s W H N in.raw out.bin [edge_index.bin [planes.bin]]\n", argv[0]); return 2; }
    const int W = atoi(argv[1]), H = atoi(argv[2]), N = atoi(argv[3]);
    FILE* fi = fopen(argv[4], "rb");
    FILE* fo = fopen(argv[5], "wb");
    FILE* fe = argc > 6 && argv[6][0] != '-' ? fopen(argv[6], "wb") : NULL;
    FILE* fp = argc > 7 ? fopen(argv[7], "wb") : NULL;
    std::vector<int16_t> G((size_t)W * H);
    std::vector<uint8_t> F((size_t)W * H);
    if (!fi || !fo) { perror("open"); return 1; }
    std::vector<unsigned char> im((size_t)W * H);
    std::vector<int32_t> out, edge((size_t)W * H);
    out.push_back(N);
    double det_ms = 0.0;
    for (int i = 0; i < N; ++i) {
        if (fread(im.data(), 1, im.size(), fi) != im.size()) { fprintf(stderr, "short read\n"); return 1; }
        sdm_host::EdgeChains ch;
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        sdm_host::DetectEdgesByED(im.data(), (size_t)W, W, H, 36, 8, ch, edge.data(), (size_t)W * 4);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        det_ms += (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
        out.push_back(ch.n_chains());
        for (int s = 0; s < ch.n_chains(); ++s) {
            out.push_back(ch.offsets[s + 1] - ch.offsets[s]);
            for (int j = ch.offsets[s]; j < ch.offsets[s + 1]; ++j) {
                out.push_back((int32_t)(ch.pixels[j] >> 16));
                out.push_back((int32_t)(ch.pixels[j] & 0xffffu));
            }
        }
        if (fe) fwrite(edge.data(), 4, edge.size(), fe);
        if (fp && W >= 5 && H >= 5) {
            sdm_host::EdPlanesHost(im.data(), (size_t)W, W, H, 36, 8, G.data(), F.data());
            fwrite(G.data(), 2, G.size(), fp);
            fwrite(F.data(), 1, F.size(), fp);
        }
    }
    fprintf(stderr, "host_ms_per_image %.4f\n", N > 0 ? det_ms / N : 0.0);
    fwrite(out.data(), sizeof(int32_t), out.size(), fo);
    fclose(fo);
    fclose(fi);
    if (fe) fclose(fe);
    if (fp) fclose(fp);
    return 0;
}
