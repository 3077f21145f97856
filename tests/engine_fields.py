"""Field lists of the two ctypes mirrors of qb_options (test helper)."""


def package_fields():
    from qutip_b200._lib import QbOptions
    return [f[0] for f in QbOptions._fields_]


def emulator_fields():
    from _emul import QbOptions
    return [f[0] for f in QbOptions._fields_]
