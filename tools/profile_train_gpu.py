"""Where one iteration of the reference's RPODDPG.run goes when it drives the kernels (configs[2]): cProfile of 300
iterations after a 50-iteration warm-up; host-side view (the loop is host-driven, GPU kernels are asynchronous)."""
import contextlib, cProfile, io, os, pstats, sys, warnings
sys.path.insert(0, '.')
os.environ["RPO_REFERENCE_ROOT"] = os.path.join(os.getcwd(), "oracle", "_ref")
import numpy as np, torch
import bench
algo_name = sys.argv[1] if len(sys.argv) > 1 else "ddpg"
rpo, _ = bench.load_reference_copy()
import gym
import rpo_b200.algo as balgo
balgo.install(rpo)
np.random.seed(111); torch.manual_seed(123)
env = gym.make("EVOPF-v0")
algo = bench._reference_algo(torch.device("cuda"), algo_name, max_epochs=50, env=env)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    with open(os.devnull, "w") as null, contextlib.redirect_stdout(null):
        algo.run(eval=False)
        algo.max_epochs = 300
        pr = cProfile.Profile()
        torch.cuda.synchronize()
        pr.enable(); algo.run(eval=False); torch.cuda.synchronize(); pr.disable()
st = io.StringIO()
ps = pstats.Stats(pr, stream=st).sort_stats("cumulative")
ps.print_stats(45)
print(st.getvalue()[:9000])
