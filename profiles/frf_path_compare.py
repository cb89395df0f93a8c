import numpy as np, sys
for nd in (100, 150):
    a = np.load(f"gpurun_out/frfsum_mma_{nd}.npy"); b = np.load(f"gpurun_out/frfsum_recurrence_{nd}.npy")
    a4, b4 = a.reshape(4, nd), b.reshape(4, nd)
    mag_u = np.hypot(b4[0], b4[1]); mag_y = np.hypot(b4[2], b4[3])
    eu = np.hypot(a4[0]-b4[0], a4[1]-b4[1]) / mag_u; ey = np.hypot(a4[2]-b4[2], a4[3]-b4[3]) / mag_y
    print(nd, "max rel diff u", eu.max(), "y", ey.max(), "argmax", eu.argmax(), ey.argmax())
