"""CPU restatement of the Edge Drawing detector as the reference calls it - TEST INFRASTRUCTURE ONLY.

    EdgeMap* map = DetectEdgesByED(srcImg, width, height, SOBEL_OPERATOR, 36, 8, 1.0)      /root/reference/src/LineDetector.cc:855

The implementation behind that call is the closed-source Thirdparty/EDTest/EDLib.a; its published algorithm (Topal & Akinlar,
JVCIR 2012) is restated here in plain Python / numpy, and every choice the paper leaves open was fixed by comparing with the
library's own output (oracle/ed_chains.cpp links the binary where it lies; the chains it finds are committed as
tests/golden/ed_chains_small.npz and ed_chains_misc.npz): smoothing = 5x5 binomial with round-half-to-even on a replicated
border, gradient = |gx| + |gy| of Sobel with border = threshold - 1, anchors on every row / column from 2 to size - 3, routing
depth first with a rotation-symmetric order of the diagonal look-ahead, extraction with the library's clean-up rules including
its read of the previous segment's last pixel.  tests/test_edge_drawing.py checks this restatement AND the product's C++
implementation (eao-slam_b200/host/edge_drawing.h) against the golden chains: identical, pixel for pixel and in order.
Only tests/ may import this module."""
import numpy as np, sys
sys.setrecursionlimit(100000)
EV, EH = 1, 2
ANCHOR, EDGE = 254, 255
LEFT, RIGHT, UP, DOWN = 1, 2, 3, 4

def smooth(im):
    a = np.pad(im.astype(np.int64), 2, mode='edge'); H, W = im.shape; k = [1, 4, 6, 4, 1]
    row = sum(k[i] * a[:, i:i + W] for i in range(5)); s = sum(k[i] * row[i:i + H, :] for i in range(5))
    even = np.rint(s / 256.0).astype(np.int64)   # the library's (OpenCV 2.4.5 cvSmooth) 4-wide vector loop: half to even
    up = (s + 128) >> 8                          # its scalar tail over the last W % 4 columns: half up
    return np.where(np.arange(W)[None, :] >= (W & ~3), up, even)

def gradient(sm, thr):
    H, W = sm.shape
    G = np.full((H, W), thr - 1, np.int64); D = np.zeros((H, W), np.int64)
    A = sm
    com1 = A[2:, 2:] - A[:-2, :-2]; com2 = A[:-2, 2:] - A[2:, :-2]
    gx = np.abs(com1 + com2 + 2 * (A[1:-1, 2:] - A[1:-1, :-2])); gy = np.abs(com1 - com2 + 2 * (A[2:, 1:-1] - A[:-2, 1:-1]))
    G[1:-1, 1:-1] = gx + gy
    m = G[1:-1, 1:-1] >= thr
    D[1:-1, 1:-1] = np.where(m, np.where(gx >= gy, EV, EH), 0)
    return G, D

def anchors(G, D, thr, athr, scan):
    H, W = G.shape
    E = np.zeros((H, W), np.int64)
    for i in range(2, H - 2):
        start, inc = 2, 1
        if i % scan != 0:
            start, inc = scan, scan
        for j in range(start, W - 2, inc):
            if G[i, j] < thr: continue
            if D[i, j] == EV:
                if G[i, j] - G[i, j - 1] >= athr and G[i, j] - G[i, j + 1] >= athr: E[i, j] = ANCHOR
            else:
                if G[i, j] - G[i - 1, j] >= athr and G[i, j] - G[i + 1, j] >= athr: E[i, j] = ANCHOR
    return E

class Chain:
    __slots__ = ('len', 'parent', 'dir', 'children', 'start')

def longest(chains, root):
    if root == -1 or chains[root].len == 0: return 0
    l0 = longest(chains, chains[root].children[0]) if chains[root].children[0] != -1 else 0
    l1 = longest(chains, chains[root].children[1]) if chains[root].children[1] != -1 else 0
    if l0 >= l1:
        mx = l0; chains[root].children[1] = -1
    else:
        mx = l1; chains[root].children[0] = -1
    return chains[root].len + mx

def retrieve(chains, root):
    out = []
    while root != -1:
        out.append(root)
        root = chains[root].children[0] if chains[root].children[0] != -1 else chains[root].children[1]
    return out

def join(G, D, E, thr, min_path, cleanup=True, rest_min=10):
    H, W = G.shape
    ys, xs = np.nonzero(E[1:-1, 1:-1] == ANCHOR); ys += 1; xs += 1
    g = G[ys, xs]
    # counting sort as sortAnchorsByGradValue1: ascending by grad, within a grad raster-later first; iterate from the end
    order = np.lexsort((-(ys * W + xs), g))   # primary g ascending, secondary offset descending
    A = (ys * W + xs)[order]
    segments = []
    prev_last = [None]
    for k in range(len(A) - 1, -1, -1):
        i, j = divmod(int(A[k]), W)
        if E[i, j] != ANCHOR: continue
        chains = [Chain()]
        chains[0].len = 0; chains[0].parent = -1; chains[0].dir = 0; chains[0].children = [-1, -1]; chains[0].start = 0
        pixels = []
        dup = 0
        stack = []
        if D[i, j] == EV:
            stack.append((i, j, DOWN, 0)); stack.append((i, j, UP, 0))
        else:
            stack.append((i, j, RIGHT, 0)); stack.append((i, j, LEFT, 0))
        while stack:
            r, c, d, parent = stack.pop()
            if E[r, c] != EDGE: dup += 1
            ch = Chain(); ch.dir = d; ch.parent = parent; ch.children = [-1, -1]; ch.start = len(pixels); ch.len = 0
            chains.append(ch); no = len(chains) - 1
            pixels.append((r, c)); clen = 1
            horizontal = d in (LEFT, RIGHT)
            child = 0 if d in (LEFT, UP) else 1
            ended = False
            while D[r, c] == (EH if horizontal else EV):
                E[r, c] = EDGE
                if horizontal:
                    if E[r - 1, c] == ANCHOR: E[r - 1, c] = 0
                    if E[r + 1, c] == ANCHOR: E[r + 1, c] = 0
                    dc = -1 if d == LEFT else 1
                    fo = -1 if d == LEFT else 1
                    if E[r, c + dc] >= ANCHOR: c += dc
                    elif E[r + fo, c + dc] >= ANCHOR: r += fo; c += dc
                    elif E[r - fo, c + dc] >= ANCHOR: r -= fo; c += dc
                    else:
                        a_, b_, c_ = G[r - 1, c + dc], G[r, c + dc], G[r + 1, c + dc]
                        if a_ > b_:
                            if a_ > c_: r -= 1
                            else: r += 1
                        elif c_ > b_: r += 1
                        c += dc
                else:
                    if E[r, c - 1] == ANCHOR: E[r, c - 1] = 0
                    if E[r, c + 1] == ANCHOR: E[r, c + 1] = 0
                    dr = -1 if d == UP else 1
                    fo = -1 if d == UP else 1
                    if E[r + dr, c] >= ANCHOR: r += dr
                    elif E[r + dr, c + fo] >= ANCHOR: r += dr; c += fo
                    elif E[r + dr, c - fo] >= ANCHOR: r += dr; c -= fo
                    else:
                        a_, b_, c_ = G[r + dr, c - 1], G[r + dr, c], G[r + dr, c + 1]
                        if a_ > b_:
                            if a_ > c_: c -= 1
                            else: c += 1
                        elif c_ > b_: c += 1
                        r += dr
                if E[r, c] == EDGE or G[r, c] < thr:
                    if clen > 0:
                        ch.len = clen
                        chains[parent].children[child] = no
                    else:
                        chains.pop()
                    ended = True
                    break
                pixels.append((r, c)); clen += 1
            if ended: continue
            if horizontal:
                stack.append((r, c, DOWN, no)); stack.append((r, c, UP, no))
            else:
                stack.append((r, c, RIGHT, no)); stack.append((r, c, LEFT, no))
            pixels.pop(); clen -= 1
            ch.len = clen
            chains[parent].children[child] = no
        ln = len(pixels)
        if ln - dup < min_path:
            for (y, x) in pixels: E[y, x] = 0
            continue
        def cpix(cno, l): return pixels[chains[cno].start + l]
        seg = []
        tot = longest(chains, chains[0].children[1])
        if tot > 0:
            nos = retrieve(chains, chains[0].children[1])
            for cno in reversed(nos):
                if cleanup:
                    fr, fc = cpix(cno, chains[cno].len - 1)
                    idx = len(seg) - 2
                    while idx >= 0:
                        if abs(fr - seg[idx][0]) <= 1 and abs(fc - seg[idx][1]) <= 1:
                            seg.pop(); idx -= 1
                        else: break
                    lastp = seg[-1] if seg else prev_last[0]
                    if chains[cno].len > 1 and lastp is not None:
                        fr, fc = cpix(cno, chains[cno].len - 2)
                        if abs(fr - lastp[0]) <= 1 and abs(fc - lastp[1]) <= 1: chains[cno].len -= 1
                for l in range(chains[cno].len - 1, -1, -1): seg.append(cpix(cno, l))
                chains[cno].len = 0
        tot = longest(chains, chains[0].children[0])
        if tot > 1:
            nos = retrieve(chains, chains[0].children[0])
            first = nos[0]
            chains[first].start += 1; chains[first].len -= 1
            for cno in nos:
                start = 0
                if cleanup:
                    fr, fc = cpix(cno, 0)
                    idx = len(seg) - 2
                    while idx >= 0:
                        if abs(fr - seg[idx][0]) <= 1 and abs(fc - seg[idx][1]) <= 1:
                            seg.pop(); idx -= 1
                        else: break
                    lastp = seg[-1] if seg else prev_last[0]
                    if chains[cno].len > 1 and lastp is not None:
                        fr, fc = cpix(cno, 1)
                        if abs(fr - lastp[0]) <= 1 and abs(fc - lastp[1]) <= 1: start = 1
                for l in range(start, chains[cno].len): seg.append(cpix(cno, l))
                chains[cno].len = 0
        if cleanup and len(seg) > 1:
            fr, fc = seg[1]
            if abs(fr - seg[-1][0]) <= 1 and abs(fc - seg[-1][1]) <= 1: seg.pop(0)
        segments.append(seg)
        prev_last[0] = seg[-1] if seg else prev_last[0]
        for kk in range(2, len(chains)):
            if chains[kk].len < 2: continue
            tot = longest(chains, kk)
            if tot >= rest_min:
                nos = retrieve(chains, kk)
                seg = []
                for cno in nos:
                    start = 0
                    if cleanup:
                        fr, fc = cpix(cno, 0)
                        idx = len(seg) - 2
                        while idx >= 0:
                            if abs(fr - seg[idx][0]) <= 1 and abs(fc - seg[idx][1]) <= 1:
                                seg.pop(); idx -= 1
                            else: break
                        if chains[cno].len > 1:
                            lastp = seg[-1] if seg else prev_last[0]
                            fr, fc = cpix(cno, 1)
                            if lastp is not None and abs(fr - lastp[0]) <= 1 and abs(fc - lastp[1]) <= 1: start = 1
                    for l in range(start, chains[cno].len): seg.append(cpix(cno, l))
                    chains[cno].len = 0
                segments.append(seg)
                prev_last[0] = seg[-1] if seg else prev_last[0]
    return segments, E

def detect(im, thr=36, athr=8):
    """im [H, W] uint8 -> list of chains, each a list of (row, col)"""
    sm = smooth(im); G, D = gradient(sm, thr); E = anchors(G, D, thr, athr, 1)
    return join(G, D, E, thr, 10)[0]
