import sys, time, numpy as np
sys.path.insert(0, ".")
import pyinterpolate_b200 as pib
from pyinterpolate_b200.synthetic import dem_like_field
for n in (20000, 100000):
    pts = dem_like_field(n, seed=3)
    for rep in range(2):
        t0 = time.time(); r = pib.calculate_semivariance(pts, 500, 40500, direction=45, tolerance=0.2); dt = time.time() - t0
    print(f"triangle n={n}: {dt*1e3:.1f} ms, pairs counted {np.nansum(r[:,2]):.3e}")
    for rep in range(2):
        t0 = time.time(); r = pib.calculate_semivariance(pts, 500, 40500, custom_weights=np.ones(n)); dt = time.time() - t0
    print(f"weighted omni n={n}: {dt*1e3:.1f} ms")
