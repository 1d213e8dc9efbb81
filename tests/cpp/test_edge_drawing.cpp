// test_edge_drawing.cpp - runs the open Edge Drawing implementation (eao-slam_b200/host/edge_drawing.h) on raw 8-bit images
// and dumps the chains in the format of oracle/ed_chains.cpp (the same dump of the reference's closed-source EDLib.a), so that
// tests/test_edge_drawing.py can compare the two byte for byte.
//   usage: test_edge_drawing W H N in.raw out.bin [edge_index.bin [planes.bin]]      ("-" skips edge_index.bin)
//   planes.bin: per image the stage-1 planes of EdPlanesHost, G (int16 W*H) then F (uint8 W*H); for widths that are multiples
//   of four (and >= 8 x 8, the library's minimum) the device kernel k_ed_planes4 is replayed thread by thread on the CPU
//   (sdm::EdPlanes4Replay, csrc/edge_drawing_kernels.cuh) and must give the same planes: exit code 4 otherwise
//   every image is also routed by the fixed-capacity form of the routing core (sdm_host::EdRouteFixed: what one warp of
//   k_ed_route runs on the device) and must give the same chains and the same edge index: exit code 5 otherwise
//   prints the host time of the detector per image (both stages, one thread) on stderr
//   out.bin (int32): N, then per image: noSegments, per segment: noPixels, (r, c) * noPixels
#include <stdio.h>
#include <time.h>
#include <stdlib.h>

#include <vector>

#include "../../eao-slam_b200/host/edge_drawing.h"
#include "../../eao-slam_b200/csrc/edge_drawing_kernels.cuh"

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
        if (W >= 5 && H >= 5) {
            sdm_host::EdPlanesHost(im.data(), (size_t)W, W, H, 36, 8, G.data(), F.data());
            const sdm_host::EdRouteCaps caps = sdm_host::EdRouteCapsFor((size_t)W * H);
            std::vector<uint8_t> scratch(sdm_host::EdRouteScratchBytes(caps) + 16);
            uint8_t* sp = scratch.data() + ((16 - ((size_t)scratch.data() & 15)) & 15);
            std::vector<int32_t> fo_(caps.offsets), edge2((size_t)W * H, -1);
            std::vector<uint32_t> fp_(caps.out_pixels);
            int nc = -1, np = -1;
            const bool ok = sdm_host::EdRouteFixed(W, H, G.data(), F.data(), 36, sp, caps, fo_.data(), fp_.data(), edge2.data(), (size_t)W * 4, &nc, &np);
            if (!ok || nc != ch.n_chains() || np != (int)ch.pixels.size() || !std::equal(ch.offsets.begin(), ch.offsets.end(), fo_.begin()) ||
                !std::equal(ch.pixels.begin(), ch.pixels.end(), fp_.begin()) || edge2 != edge) {
                fprintf(stderr, "fixed-capacity routing differs at image %d (ok %d, chains %d / %d, pixels %d / %d)\n", i, (int)ok, nc,
                        ch.n_chains(), np, (int)ch.pixels.size());
                return 5;
            }
        }
        if (fp && W >= 5 && H >= 5) {
            sdm_host::EdPlanesHost(im.data(), (size_t)W, W, H, 36, 8, G.data(), F.data());
            fwrite(G.data(), 2, G.size(), fp);
            fwrite(F.data(), 1, F.size(), fp);
            if ((W & 3) == 0 && W >= 8 && H >= 8) {
                std::vector<int16_t> G4((size_t)W * H, (int16_t)-1);
                std::vector<uint8_t> F4((size_t)W * H, (uint8_t)0xee);
                sdm::EdPlanes4Replay(im.data(), W, H, 36, 8, G4.data(), F4.data());
                if (G4 != G || F4 != F) {
                    size_t k = 0;
                    while (k < G.size() && G4[k] == G[k] && F4[k] == F[k]) ++k;
                    fprintf(stderr, "k_ed_planes4 replay differs at image %d, pixel (%d, %d): G %d / %d, F %02x / %02x\n", i,
                            (int)(k / W), (int)(k % W), (int)G4[k], (int)G[k], (unsigned)F4[k], (unsigned)F[k]);
                    return 4;
                }
            }
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
