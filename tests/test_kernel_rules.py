"""CPU: the integer / indexing rules the kernels rely on, restated in Python and checked exhaustively on small
cases (regression guards: the kernels themselves are checked bit-for-bit on the GPU in test_gpu_parity.py)."""
from math import gcd

import numpy as np


def _window_sum_loop(m, N, hop, T, w2):
    """The kernels' edge loop: frames ascending, float32 accumulation (stft_lowpass.cu / stft_splice.cu)."""
    fa = (m - N) // hop + 1 if m - N >= 0 else 0
    fb = min(T - 1, m // hop)
    ws = np.float32(0)
    for f in range(fa, fb + 1):
        ws = np.float32(ws + w2[m - f * hop])
    return ws


def test_interior_window_sum_table_rule():
    """K4 / K6: for a sample m with m >= N - hop and m // hop <= T - 1 the overlap-added window^2 equals the plan's
    table entry m % hop (same additions, same order); the 32-bit counters reproduce m // hop and m % hop."""
    N = 64
    rng = np.random.default_rng(0)
    for hop in (7, 16, 21, 33, 64):
        w2 = rng.random(N).astype(np.float32)
        table = np.zeros(hop, np.float32)
        for r in range(hop):
            ws = np.float32(0)
            for j in range((N - 1 - r) // hop, -1, -1):
                ws = np.float32(ws + w2[r + j * hop])
            table[r] = ws
        for L in (N // 2 + 1, 100, 257, 1000):
            T = L // hop + 1
            chunk_hops, threads = 5, 8
            for bx in range((L + chunk_hops * hop - 1) // (chunk_hops * hop)):
                n0 = bx * chunk_hops * hop
                n1 = min(L, n0 + chunk_hops * hop)
                m0, span = n0 + N // 2, n1 - n0
                for tid in range(threads):
                    t = N // 2 + tid
                    q, r = t // hop, t % hop
                    for i in range(tid, span, threads):
                        m = m0 + i
                        assert bx * chunk_hops + q == m // hop and r == m % hop
                        interior = m >= N - hop and bx * chunk_hops + q <= T - 1
                        if interior:
                            assert table[r] == _window_sum_loop(m, N, hop, T, w2), (hop, L, m)
                        r += threads
                        while r >= hop:
                            r -= hop
                            q += 1


def test_tmem_sample_ring_chunk_mapping():
    """K1 (hop 512 = 4 blocks of 128): block r of frame f is global block 4f - 8 + r and lives in ring chunk
    (f + 2 + r // 4) % 4 -- the same chunk whichever frame asks for it; frame f + 1 replaces exactly the chunk that
    held frame f's oldest four blocks."""
    home = {}
    for f in range(2, 200):
        for r in range(16):
            block, chunk = 4 * f - 8 + r, (f + 2 + r // 4) % 4
            assert home.setdefault(block, chunk) == chunk
        new_chunk = (f + 1 + 2 + 3) % 4          # chunk written for frame f + 1 (its r = 12..15)
        assert new_chunk == (f + 2 + 0) % 4      # = the chunk of frame f's r = 0..3


def test_resampler_span_bound_and_staged_indices():
    """K3: a CTA's outputs [jb, jb + TP*R) touch inputs newest(jb) - (K-1) .. newest(jb + TP*R - 1); the staged span
    (TP*R - 1) * down // up + 2 + K covers them, and output j0 + m*TP needs newest(j0) + m*step."""
    R = 8
    for up, down in ((160, 147), (147, 160), (441, 160), (80, 147), (3, 1), (1, 2), (3, 2), (5, 7)):
        g = gcd(up, down)
        up, down = up // g, down // g
        n_taps = 2 * 10 * max(up, down) + 1
        half = (n_taps - 1) // 2
        n_pre_pad = down - half % down
        n_pre_remove = (half + n_pre_pad) // down
        K = (n_taps + up - 1) // up
        TP = up * ((256 + up - 1) // up)
        step = (TP // up) * down
        span = (TP * R - 1) * down // up + 2 + K

        def newest(j):
            return ((j + n_pre_remove) * down - n_pre_pad) // up   # Python floor division = the kernel's fix-up

        for block in (0, 1, 7):
            jb = block * TP * R
            i_base = newest(jb) - (K - 1)
            assert newest(jb + TP * R - 1) - i_base < span
            for tid in (0, 1, TP // 2, TP - 1):
                j0 = jb + tid
                p0 = newest(j0) - i_base
                assert p0 - (K - 1) >= 0
                for m in range(R):
                    assert newest(j0 + m * TP) == newest(j0) + m * step
                    assert p0 + m * step < span


def test_pass3_hermitian_pairing_covers_every_bin_once():
    """K1 / K4 / K6 (k1_map.cuh): thread t owns last-pass butterflies ia, ib; bins k_low(i) + 256 q, q = 0..3, of
    both, plus bin 1024 on the special thread, are exactly the bins 0..1024, and the partner of every emitted bin
    sits in the thread's other butterfly at q' = 7 - q (or (8 - q) % 8 inside butterfly 0)."""
    def klow(i):
        return (i >> 4) + ((i & 15) << 4)

    def butterflies(t):
        if t < 120:
            return 16 + t, 255 - t
        if t < 127:
            return t - 119, 16 - (t - 119)
        return 0, 8

    seen = []
    for t in range(128):
        ia, ib = butterflies(t)
        ka, kb = klow(ia), klow(ib)
        for q in range(4):
            for (k, own, other) in ((ka + 256 * q, ia, ib), (kb + 256 * q, ib, ia)):
                seen.append(k)
                partner = (2048 - k) % 2048
                if t == 127:  # butterfly 0 pairs inside itself, butterfly 8 inside itself
                    pq = (8 - q) % 8 if own == 0 else 7 - q
                    assert partner == klow(own) + 256 * pq
                else:
                    assert partner == klow(other) + 256 * (7 - q)
        if t == 127:
            seen.append(1024)
    assert sorted(seen) == list(range(1025))
