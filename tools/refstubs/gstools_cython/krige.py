"""``gstools_cython.krige`` stand-in bound to the CPU oracle (imported at
/root/reference/src/gstools/krige/base.py:16-19)."""
import oracle as _oracle


def calc_field_krige_and_variance(krig_mat, krig_vecs, cond, num_threads=None):
    return _oracle.calc_field_krige_and_variance(krig_mat, krig_vecs, cond, num_threads)


def calc_field_krige(krig_mat, krig_vecs, cond, num_threads=None):
    return _oracle.calc_field_krige(krig_mat, krig_vecs, cond, num_threads)
