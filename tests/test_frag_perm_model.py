"""Lane-by-lane CPU model of the fragment mapping of the TMA-fed DMMA kernels (csrc/it_gemm_tma.cuh): 128-byte-swizzled
shared-memory rows, 128-bit fragment loads, mma.sync.m8n8k4.f64 operand / accumulator ownership.  It checks, for the default
mapping and for the permuted one (LOWDIN_IT_OPT_FRAG_PERM, frag_row), that every accumulator element lands on the output
element the epilogue indexes, and counts the shared-memory bank conflicts of a fragment load per quarter-warp phase.
(No kernel runs here; the GPU tests of the variant are in tests/test_gpu_variants.py.)"""
import numpy as np
import pytest


def swizzled(tile):
    """[8 rows][16 doubles] -> storage [row][16-byte chunk][2]; TMA SWIZZLE_128B: chunk ^= row & 7"""
    P = np.zeros((8, 8, 2))
    for r in range(8):
        for c in range(8):
            P[r, c ^ (r & 7)] = tile[r, 2 * c:2 * c + 2]
    return P


def rho(g, perm):
    return ((g >> 1) | ((g & 1) << 2)) if perm else g


@pytest.mark.parametrize("perm", [False, True])
def test_fragment_mapping_and_bank_conflicts(perm):
    rng = np.random.default_rng(1)
    A, B = rng.standard_normal((8, 16)), rng.standard_normal((8, 16))      # one 8 x 8 output tile, one k-tile of 16
    PA, PB = swizzled(A), swizzled(B)
    acc = np.zeros((32, 2))
    conflicts = 0
    for h in range(2):                                                      # the two k-halves of a k-tile
        a, b = np.zeros((32, 2)), np.zeros((32, 2))
        for phase in range(4):                                              # a 128-bit load is served one quarter-warp per cycle
            chunks = []
            for lane in range(8 * phase, 8 * phase + 8):
                g, t = lane >> 2, lane & 3
                r = rho(g, perm)
                ch = (t ^ r) ^ (4 if h else 0)                              # off0 = (tig ^ row) << 4; k-half 1 is off0 ^ 64
                a[lane], b[lane] = PA[r, ch], PB[r, ch]
                chunks.append(ch)                                           # rows are 128 bytes apart: same banks for the same chunk
            conflicts += len(chunks) - len(set(chunks))
        for e in range(2):                                                  # .x feeds one DMMA, .y the next
            # mma.m8n8k4: lane (g,t) supplies A[g][t] and B[t][g]; owns D[g][2t], D[g][2t+1]
            D = np.array([[sum(a[g * 4 + t, e] * b[c * 4 + t, e] for t in range(4)) for c in range(8)] for g in range(8)])
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                acc[lane] += D[g, 2 * t], D[g, 2 * t + 1]
    ref = A @ B.T
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        m = rho(g, perm)
        n0, n1 = (t, t + 4) if perm else (2 * t, 2 * t + 1)                 # frag_col0 / frag_col1
        assert abs(acc[lane, 0] - ref[m, n0]) < 1e-12 and abs(acc[lane, 1] - ref[m, n1]) < 1e-12
    # default mapping: the two rows of every quarter-warp share a half line -> 4 of 8 lanes conflict in each of 8 phases
    assert conflicts == (0 if perm else 32)
