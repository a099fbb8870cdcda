"""Generates tests/golden/sh_common_golden.npz from the REFERENCE's own shaders/sh_common.h
(compiled by oracle/build_oracle.py into oracle/_ref/libvlb_refsh.so). Run in the container
where /root/reference is mounted; the committed .npz lets the same check run on the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_oracle, oracle_api as oa  # noqa: E402


def main():
    build_oracle.build_ref(force=True)
    R = oa.ref_lib()
    assert R is not None, "reference header not available"
    rng = np.random.default_rng(20261017)
    d = rng.normal(size=(512, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    # a few exact axis / diagonal directions
    d[:6] = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]], np.float32)
    basis = np.zeros((len(d), 25), np.float32)
    for n in range(len(d)):
        i = 0
        for l in range(5):
            for m in range(-l, l + 1):
                basis[n, i] = R.ref_SH(l, m, float(d[n, 0]), float(d[n, 1]), float(d[n, 2]))
                i += 1
    sizes = np.array([[1, 1], [7, 13], [64, 32], [3141, 1000], [2048, 1024]], np.int32)
    phis, thetas = [], []
    for w, h in sizes:
        phis.append(np.array([R.ref_x2phi(x, int(w)) for x in range(min(w, 64))], np.float32))
        thetas.append(np.array([R.ref_y2theta(y, int(h)) for y in range(min(h, 64))], np.float32))
    ang = rng.uniform(0, 1, (256, 2)).astype(np.float32) * np.array([2 * np.pi, np.pi], np.float32)
    vec = np.zeros((256, 3), np.float32)
    for n in range(256):
        import ctypes
        R.ref_toVector(float(ang[n, 0]), float(ang[n, 1]), vec[n].ctypes.data_as(ctypes.c_void_p))
    out = os.path.join(ROOT, "tests", "golden", "sh_common_golden.npz")
    np.savez_compressed(out, dirs=d, basis=basis, sizes=sizes, angles=ang, to_vector=vec, pi=np.float32(R.ref_PI()),
                        **{"phi_%d" % i: p for i, p in enumerate(phis)}, **{"theta_%d" % i: t for i, t in enumerate(thetas)})
    print("wrote", out)


if __name__ == "__main__":
    main()
