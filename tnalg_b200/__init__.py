"""tnalg_b200: B200-native (sm_100a) finite-size DMRG hot path of ranshiju/T-Nalg behind the reference's Python API.

    from tnalg_b200 import Parameters, DMRG_anyH
    para = Parameters.generate_parameters_dmrg('chain'); ...
    ob, A, info, para = DMRG_anyH.dmrg_finite_size(para)

tnalg_b200/dropin/ holds flat modules named like the reference's (MPSClass, DMRG_anyH, Parameters, ...), so putting
that directory on sys.path in place of the reference tree keeps existing scripts and `.pr` pickles working.
"""
__version__ = '0.1.0'
