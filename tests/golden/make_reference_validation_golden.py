"""Extract the reference's OWN regression data for the laser-wake benchmark into a small fixture.

    python tests/golden/make_reference_validation_golden.py      (needs /root/reference)

validation/references/tst3d_s_o2_laser_wake_yee_vay.py.txt is the pickle Smilei's validation script compares
against (validation/analyses/validate_tst3d_s_o2_laser_wake_yee_vay.py): Ey on the central axis, every 4th of
the 512 probe points, at timesteps 300 and 1000, absolute tolerance 0.01.
"""
import os
import pickle

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/validation/references/tst3d_s_o2_laser_wake_yee_vay.py.txt"

if __name__ == "__main__":
    with open(SRC, "rb") as f:
        d = pickle.load(f, encoding="latin1")
    out = {"Ey_axis_300": np.asarray(d["Field Ey on central axis timestep 300"], dtype=np.float64),
           "Ey_axis_1000": np.asarray(d["Field Ey on central axis timestep 1000"], dtype=np.float64),
           "tolerance": np.float64(0.01)}
    np.savez_compressed(os.path.join(HERE, "ref_validation_laser_wake_vay.npz"), **out)
    print({k: (v.shape, float(np.max(np.abs(v)))) for k, v in out.items()})
