"""``gstools_cython.krige`` stand-in (imported at krige/base.py:16-19)."""
import numpy as np


def calc_field_krige_and_variance(krig_mat, krig_vecs, cond, num_threads=None):
    kv = np.asarray(krig_vecs)
    mk = np.asarray(krig_mat) @ kv
    return np.asarray(cond) @ mk, np.einsum("ij,ij->j", kv, mk)


def calc_field_krige(krig_mat, krig_vecs, cond, num_threads=None):
    return np.asarray(cond) @ (np.asarray(krig_mat) @ np.asarray(krig_vecs))
