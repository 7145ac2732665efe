"""One C4 rollout batch (train_metabo_gps.py environment with T = 100, M = 2000 Sobol candidates, Matern-5/2, 2 x 20 policy) for an ncu
launch list: MBO_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum ... python tools/c4_rollout_ncu.py [episodes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import run_configs as rc
n_ep = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
spec = dict(rc.GPS, T=100)
br = rc.recorder(spec, "weights_gps_1161.npz", 10, n_ep, 100, True, "fp32")
for i in range(2):
    t = br.record_batch(1.0, 1.0)
    print("batch", i, "t_batch %.4f s  %.0f env-steps/s" % (t, n_ep * 100 / t), flush=True)
