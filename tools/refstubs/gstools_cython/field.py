"""``gstools_cython.field`` stand-in bound to the CPU oracle (imported at
/root/reference/src/gstools/field/generator.py:22-24)."""
import oracle as _oracle


def summate(cov_samples, z_1, z_2, pos, num_threads=None):
    return _oracle.summate(cov_samples, z_1, z_2, pos, num_threads)


def summate_incompr(cov_samples, z_1, z_2, pos, num_threads=None):
    return _oracle.summate_incompr(cov_samples, z_1, z_2, pos, num_threads)


def summate_fourier(spectrum_factor, modes, z_1, z_2, pos, num_threads=None):
    return _oracle.summate_fourier(spectrum_factor, modes, z_1, z_2, pos, num_threads)
