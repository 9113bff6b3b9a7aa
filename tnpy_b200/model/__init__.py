from tnpy_b200.model.model_1d import Model1D
from tnpy_b200.model.spin_chains import XXZ, DimerXXZ, RandomHeisenberg, Thirring, TotalSz, TransverseIsing

__all__ = ["Model1D", "TotalSz", "XXZ", "Thirring", "RandomHeisenberg", "DimerXXZ", "TransverseIsing"]
