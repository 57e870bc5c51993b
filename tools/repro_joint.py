import sys, numpy as np, torch
sys.path.insert(0, ".")
from optistate_b200 import Kalman_Filter
kf = Kalman_Filter(); kf.x = kf.x.copy()
p = np.array([0.2, 0.15, -0.28, 0.2, -0.15, -0.28, -0.2, 0.15, -0.28, -0.2, -0.15, -0.28]).reshape(12, 1)
f = np.zeros((12, 1)); f[2] = 43.0; f[11] = 43.0
kf.predict(p, f); print("predict ok", kf.P_trace)
kf.update(); print("update ok", kf.K_gain)
