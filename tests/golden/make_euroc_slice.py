"""Cuts the rows of the reference's EuRoC ground-truth file that its own bias-estimation run reads
(cpp/tests/imu_test.cpp:704-760, 885-945: samples 2000.. until 30 keyframes 0.5 s apart exist) into a small fixture.

Run HERE (the container that has /root/reference); the GPU box only sees the committed .npz.
    python tests/golden/make_euroc_slice.py
Stored: the raw columns timestamp [ns, int64], p (3), q (w, x, y, z), v (3) — what read_line_euroc (imu_test.cpp:9-39) keeps.
"""
import os

import numpy as np

SRC = "/root/reference/cpp/tests/euroc_gt.csv"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "euroc_gt_slice.npz")
IDX_START = 2000       # imu_test.cpp:888
N_KF = 30              # :939
DT_KF = 0.5            # :889


def main():
    ts, val = [], []
    with open(SRC) as f:
        next(f)                                 # header (:728)
        next(f)                                 # the first data row is read and then overwritten (:731-739)
        for line in f:
            c = line.strip().split(",")
            ts.append(int(c[0]))
            val.append([float(x) for x in c[1:11]])
    ts = np.asarray(ts, dtype=np.int64)
    val = np.asarray(val)
    # sample k of meas_vec/pose_vec/ts_vec is row k here, differentiated against row k + 1 (:741-757)
    t = ts.astype(np.float64) * 1e-9
    n_kf, last, k = 1, t[IDX_START], IDX_START
    while n_kf < N_KF:
        k += 1
        if t[k] - last > DT_KF:
            n_kf, last = n_kf + 1, t[k]
    rows = slice(IDX_START, k + 2)
    np.savez_compressed(OUT, first_sample=np.int64(IDX_START), timestamp_ns=ts[rows], p=val[rows, 0:3], q_wxyz=val[rows, 3:7], v=val[rows, 7:10])
    print(OUT, ts[rows].size, "rows", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
