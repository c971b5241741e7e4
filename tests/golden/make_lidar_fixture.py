"""Extract frame 0 of the reference's lidar log (data/double/train.dat, parsed exactly as
test/gtest/test_lidar_gp_2d.cpp:82-104: int32 numel, numel x f64 angles, numel x f64 ranges,
uint64 pose_size, pose_size x f64 pose) into a small fixture that travels to the GPU box.

Run in the build container only (needs /root/reference):  python tests/golden/make_lidar_fixture.py
"""
import os
import struct

import numpy as np

SRC = "/root/reference/data/double/train.dat"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lidar_train_frames.npz")


def frames(path):
    data = open(path, "rb").read()
    off = 0
    while off < len(data):
        (numel,) = struct.unpack_from("<i", data, off)
        off += 4
        angles = np.frombuffer(data, dtype="<f8", count=numel, offset=off)
        off += 8 * numel
        ranges = np.frombuffer(data, dtype="<f8", count=numel, offset=off)
        off += 8 * numel
        (pose_size,) = struct.unpack_from("<Q", data, off)
        off += 8
        pose = np.frombuffer(data, dtype="<f8", count=pose_size, offset=off)
        off += 8 * pose_size
        yield angles.copy(), ranges.copy(), pose.copy()


if __name__ == "__main__":
    fr = list(frames(SRC))
    keep = [0, 13, 27]
    np.savez_compressed(OUT, frame_ids=np.array(keep), angles=np.stack([fr[i][0] for i in keep]), ranges=np.stack([fr[i][1] for i in keep]), poses=np.stack([fr[i][2] for i in keep]))
    print(len(fr), "frames;", fr[0][0].shape, "beams; wrote", OUT)
