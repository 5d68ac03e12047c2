"""tests/golden/oracle_thermal_short_curves.npz: Ukin and Uelm every 10 steps of tst3d_v_o2_thermal_plasma_short
(2001 steps, the reference's particle streams of seed 0) computed by the CPU ORACLE through the repository's driver
(Simulation on tests/oracle_patch.OraclePatch).  About 16 minutes on 8 cores.

    python tests/golden/make_oracle_thermal_short.py

tests/test_gpu_simulation.py::test_reference_validation_thermal_plasma_short holds the GPU run to these curves:
a full-length trajectory parity check next to the statistical one against the reference's stored curves.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

if __name__ == "__main__":
    from oracle_patch import OraclePatch
    from smilei_b200.simulation import Simulation
    from test_reference_streams import thermal_short
    sim = Simulation(thermal_short(), patch_factory=OraclePatch)
    sim.create_particles(reference_streams=True)
    uk, ue = sim.scalars()
    K, E = [float(uk.sum())], [ue]
    for s in range(200):
        for _, k, e in sim.run(10, scalars_every=10):
            K.append(float(k.sum()))
            E.append(e)
    np.savez(os.path.join(HERE, "oracle_thermal_short_curves.npz"), ukin=np.asarray(K), uelm=np.asarray(E))
