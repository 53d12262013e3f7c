import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import _lib
import numpy as np
h = _lib.Handle(0)
Z = []
for seed in range(100, 140):
    m = h.normal_moments(seed, 400_000_000, 32)
    N = m[0]
    mean, m2, m3, m4, cross = (m[k] / N for k in (1, 2, 3, 4, 5))
    Z.append([mean * math.sqrt(N), (m2 - 1) / math.sqrt(2 / N), m3 / math.sqrt(15 / N), (m4 - 3) / math.sqrt(96 / N), 2 * cross * math.sqrt(N / 2)])
Z = np.array(Z)
print("40 seeds x 1.024e11 draws; z-scores of [mean, m2, m3, m4, pair-corr]")
print("mean of z :", np.round(Z.mean(0), 2), " (sd of mean 0.16)")
print("std  of z :", np.round(Z.std(0, ddof=1), 2))
print("max |z|   :", np.round(np.abs(Z).max(0), 2))
N = 1.024e11
print("implied bias: mean %.2e  var %.2e  m4 %.2e" % (Z[:,0].mean()/math.sqrt(N), Z[:,1].mean()*math.sqrt(2/N), Z[:,3].mean()*math.sqrt(96/N)))
