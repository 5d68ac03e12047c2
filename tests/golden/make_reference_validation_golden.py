"""Extract the reference's OWN regression data for the laser-wake benchmark into a small fixture.

    python tests/golden/make_reference_validation_golden.py      (needs /root/reference)

validation/references/tst3d_*_laser_wake_*.py.txt are the pickles Smilei's validation scripts compare against
(validation/analyses/validate_tst3d_s_o2_laser_wake_yee_vay.py and siblings): Ey on the central axis, every 4th
of the 512 probe points, at timesteps 300 and 1000, absolute tolerance 0.01.
"""
import os
import pickle

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFS = "/root/reference/validation/references/"
# the four laser-wake benchmarks share the analysis (validate_tst3d_*_laser_wake_*.py): vay / higueracary with
# oblique absorption vectors, boris (vectorised build, default vectors), boris at interpolation order 4
CASES = {"vay": "tst3d_s_o2_laser_wake_yee_vay", "higueracary": "tst3d_s_o2_laser_wake_yee_higuera",
         "boris": "tst3d_v_o2_laser_wake_yee_boris", "boris_o4": "tst3d_v_o4_laser_wake_boris"}

if __name__ == "__main__":
    out = {"tolerance": np.float64(0.01)}
    for tag, name in CASES.items():
        with open(REFS + name + ".py.txt", "rb") as f:
            d = pickle.load(f, encoding="latin1")
        out[tag + "_300"] = np.asarray(d["Field Ey on central axis timestep 300"], dtype=np.float64)
        out[tag + "_1000"] = np.asarray(d["Field Ey on central axis timestep 1000"], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_validation_laser_wake.npz"), **out)
    print({k: (np.shape(v), float(np.max(np.abs(v)))) for k, v in out.items()})
    # tst3d_00_em_propagation: a Gaussian beam at oblique incidence through Silver-Mueller sides, no plasma;
    # validate_tst3d_00_em_propagation.py checks probes of Ey with tolerance 0.01
    with open(REFS + "tst3d_00_em_propagation.py.txt", "rb") as f:
        d = pickle.load(f, encoding="latin1")
    em = {"probe0_Ey_vs_time": np.asarray(d["0-D probe Ey vs time"], dtype=np.float64),
          "probe1_Ey": np.asarray(d["1-D probe Ey at last iteration"], dtype=np.float64),
          "probe2_Ey": np.asarray(d["2-D probe Ey at last iteration"], dtype=np.float64),
          "tolerance": np.float64(0.01)}
    np.savez_compressed(os.path.join(HERE, "ref_validation_em_propagation.npz"), **em)
    print({k: (np.shape(v), float(np.max(np.abs(v)))) for k, v in em.items()})
    # tst3d_v_o2_thermal_plasma_short (== tst3d_gpu_o2_thermal_plasma_short, the same data): energy curves every 10
    # steps over 2001 steps, validate_tst3d_v_o2_thermal_plasma_short.py: Ukin/avg, Uelm/avg, Utot/avg with
    # tolerances 1e-3, 0.02, 1e-3
    with open(REFS + "tst3d_v_o2_thermal_plasma_short.py.txt", "rb") as f:
        d = pickle.load(f, encoding="latin1")
    with open(REFS + "tst3d_gpu_o2_thermal_plasma_short.py.txt", "rb") as f:
        dg = pickle.load(f, encoding="latin1")
    th = {"ukin": np.asarray(d["Ukinetic energy evolution: "], dtype=np.float64),
          "uelm": np.asarray(d["Uelectromag evolution: "], dtype=np.float64),
          "utot": np.asarray(d["Total energy evolution: "], dtype=np.float64),
          "tolerance": np.asarray([1e-3, 0.02, 1e-3])}
    assert all(np.array_equal(np.asarray(dg[k]), np.asarray(d[k])) for k in d)
    np.savez_compressed(os.path.join(HERE, "ref_validation_thermal_plasma_short.npz"), **th)
    print({k: (np.shape(v), float(np.max(np.abs(v)))) for k, v in th.items()})
    # tst3d_v_o2_thermal_plasma_medium (== tst3d_gpu_o2_thermal_plasma_medium): 51 samples, tolerance 1e-3 on all three
    with open(REFS + "tst3d_v_o2_thermal_plasma_medium.py.txt", "rb") as f:
        d = pickle.load(f, encoding="latin1")
    tm = {"ukin": np.asarray(d["Ukinetic energy evolution: "], dtype=np.float64),
          "uelm": np.asarray(d["Uelectromag evolution: "], dtype=np.float64),
          "utot": np.asarray(d["Total energy evolution: "], dtype=np.float64),
          "tolerance": np.asarray([1e-3, 1e-3, 1e-3])}
    np.savez_compressed(os.path.join(HERE, "ref_validation_thermal_plasma_medium.npz"), **tm)
