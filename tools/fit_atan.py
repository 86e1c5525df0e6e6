"""Coefficients of atan2_xpos in csrc/geometry.cu: atan(z) = z * P(z^2) on [0, 1], P of degree 8 fitted at Chebyshev
nodes, error checked in fp32 arithmetic against float64 (prints 1.09e-07)."""
import numpy as np
from numpy.polynomial import chebyshev as C

k = np.arange(400)
u = 0.5 * (1 + np.cos(np.pi * (k + 0.5) / 400))
z = np.sqrt(u)
coef = C.Chebyshev.fit(u, np.arctan(z) / z, 8, domain=[0, 1]).convert(kind=np.polynomial.Polynomial).coef
z = np.linspace(0, 1, 200001).astype(np.float32)
uu = (z * z).astype(np.float32)
p = np.float32(coef[-1])
for a in coef[-2::-1]:
    p = (p * uu + np.float32(a)).astype(np.float32)
print([float(np.float32(a)) for a in coef])
print(np.abs((z * p).astype(np.float64) - np.arctan(z.astype(np.float64))).max())
