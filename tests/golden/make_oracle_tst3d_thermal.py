"""tests/golden/oracle_tst3d_thermal_o{2,4}.npz: per-step Ukin (per species) and Uelm of the reference benchmarks
tst3d_01_thermal_plasma (order 2, BASELINE.json configs[0]) and tst3d_v_o4_thermal_plasma (order 4, configs[2] at the
reference's own size), all 163 steps from the reference's particle streams, computed by the CPU ORACLE through the
repository's driver.  Minutes on 8 cores.

    python tests/golden/make_oracle_tst3d_thermal.py [2|4]
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
    from test_reference_streams import tst3d_thermal
    for order in ([int(sys.argv[1])] if len(sys.argv) > 1 else [2, 4]):
        params = tst3d_thermal(order)
        sim = Simulation(params, patch_factory=OraclePatch)
        sim.create_particles(reference_streams=True)
        uk, ue = sim.scalars()
        K, E = [uk.copy()], [ue]
        for _, k, e in sim.run(params.n_time, scalars_every=1):
            K.append(k.copy())
            E.append(e)
        np.savez(os.path.join(HERE, f"oracle_tst3d_thermal_o{order}.npz"), ukin=np.asarray(K), uelm=np.asarray(E),
                 n_particles=np.asarray(sim.n_particles()))
        print(order, params.n_time, sim.n_particles(), K[0], K[-1], E[-1], flush=True)
