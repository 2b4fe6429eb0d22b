"""``gstools_cython.field`` stand-in bound to the CPU oracle (imported at
/root/reference/src/gstools/field/generator.py:22-24)."""
import numpy as np

import oracle as _oracle


def summate(cov_samples, z_1, z_2, pos, num_threads=None):
    return _oracle.summate(cov_samples, z_1, z_2, pos, num_threads)


def summate_incompr(cov_samples, z_1, z_2, pos, num_threads=None):
    return _oracle.summate_incompr(cov_samples, z_1, z_2, pos, num_threads)


def summate_fourier(spectrum_factor, modes, z_1, z_2, pos, num_threads=None):
    # generator.py:659-691: sum_j sqrt-spectrum_j (z1_j cos(k_j.x) + z2_j sin(k_j.x))
    modes = np.asarray(modes, dtype=np.double)
    pos = np.asarray(pos, dtype=np.double)
    out = np.zeros(pos.shape[1])
    step = max(1, 2_000_000 // max(1, modes.shape[1]))
    for a in range(0, pos.shape[1], step):
        phase = pos[:, a:a + step].T @ modes
        out[a:a + step] = np.cos(phase) @ (spectrum_factor * z_1) + np.sin(phase) @ (
            spectrum_factor * z_2)
    return out
