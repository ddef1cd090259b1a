"""CPU: the write-ownership rule of the symmetric distance epilogue (csrc/gemm_tc.cu EpiDistSym + the tile skip in
gemm_tc.cuh), replayed in numpy: over all computed tiles, warps, 32-column chunks and lanes every element of the
N x N matrix must be written exactly once, ragged edges included (N not a multiple of 32 / 128 / 256)."""
import numpy as np
import pytest

BM = 128


def _replay(n, bn):
    writes = np.zeros((n, n), dtype=np.int32)
    m_blocks, n_blocks = -(-n // BM), -(-n // bn)
    for m_blk in range(m_blocks):
        for n_blk in range(n_blocks):
            if (n_blk + 1) * bn <= m_blk * BM:                     # EpiDistSym::skip_tile
                continue
            for q in range(4):                                      # TMEM lane quarters = warps of 32 rows
                r0 = m_blk * BM + q * 32
                for c in range(bn // 32):
                    col0 = n_blk * bn + c * 32
                    if col0 >= n:
                        continue                                    # kernel: epi only if col0 < N
                    ncols = min(32, n - col0)
                    for lane in range(32):
                        row = r0 + lane
                        if row >= n:
                            continue                                # kernel: epi only if row < M
                        if col0 < (row & ~31):
                            continue                                # below the diagonal: arrives as a mirror
                        if col0 > (row & ~31) and ncols == 32:
                            writes[row, col0:col0 + 32] += 1
                            writes[col0:col0 + 32, row] += 1
                        else:
                            for j in range(col0, col0 + ncols):
                                if j >= row:
                                    writes[row, j] += 1
                                if j > row:
                                    writes[j, row] += 1
    return writes


@pytest.mark.parametrize("n,bn", [(256, 256), (300, 256), (383, 256), (416, 256), (700, 256), (130, 128), (255, 128)])
def test_every_element_is_written_exactly_once(n, bn):
    w = _replay(n, bn)
    assert w.min() == 1 and w.max() == 1, (int(w.min()), int(w.max()), np.argwhere(w != 1)[:5])
