"""Conversions between the engine's packed parity T-matrix layout (include/hydrotm.h) and the oracle's
full four-block layout (oracle/sl_oracle.c header)."""
import numpy as np


def packed_size(nmax):
    return 2 * (nmax * nmax + nmax * (nmax + 1) * (2 * nmax + 1) // 6)


def packed_to_full(tp, nmax):
    """packed [m][c][k'][k] -> full [m][blk(11,12,21,22)][k][k'] (complex128 1-D arrays)."""
    out = []
    off = 0
    for m in range(nmax + 1):
        nmin = max(m, 1)
        nm = nmax - nmin + 1
        blk = np.asarray(tp[off:off + 2 * nm * nm]).reshape(2, nm, nm)  # [c][k'][k]
        off += 2 * nm * nm
        full = np.zeros((4, nm, nm), dtype=np.complex128)  # [blk][k][k']
        for k in range(nm):
            n = nmin + k
            for kp in range(nm):
                npr = nmin + kp
                for tau in (1, 2):
                    c = (n + 1) % 2 if tau == 1 else n % 2
                    taup = 1 if (npr + 1) % 2 == c else 2
                    full[(tau - 1) * 2 + (taup - 1), k, kp] = blk[c, kp, k]
        out.append(full.reshape(-1))
    return np.concatenate(out)
