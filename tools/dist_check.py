"""Multi-GPU check (run under torchrun on >= 2 GPUs): distributed one-shot conditioning == single-GPU result.
   torchrun --nproc-per-node 2 tools/dist_check.py [npde] [nbc_edge]"""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
import linpde_gp_b200 as lg
from linpde_gp_b200.linfuncops import diffops
from linpde_gp_b200.randprocs import covfuncs
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
npde = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
nbc = int(sys.argv[2]) if len(sys.argv) > 2 else 129
prob = bench.make_problem(npde, nbc, 32)
k = 4.0 * covfuncs.TensorProduct(covfuncs.Matern((), nu=2.5, lengthscales=prob["ell"]), covfuncs.Matern((), nu=2.5, lengthscales=prob["ell"]))
prior = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), k)
batches = [(Yb, Xb) for Xb, Yb in zip(prob["edges"], prob["Y_bc"])] + [(prob["Y_pde"], prob["X_pde"], -1.0 * diffops.Laplacian((2,)))]
torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
post = lg.ConditionalGaussianProcess.from_observation_batches(prior, batches, nb=512)
torch.cuda.synchronize(); dist.barrier(); t1 = time.time()
m_d, v_d = post.mean(prob["Xt"]), post.var(prob["Xt"])
# single-GPU reference on every rank: temporarily pretend there is no process group
post1 = prior
for b in batches:
    post1 = post1.condition_on_observations(b[0], X=b[1], L=b[2] if len(b) > 2 else None)
m_1, v_1 = post1.mean(prob["Xt"]), post1.var(prob["Xt"])
sc = max(np.max(np.abs(v_1)), np.max(np.abs(m_1)))
err = max(np.max(np.abs(m_d - m_1)), np.max(np.abs(v_d - v_1))) / sc
print(f"rank {rank}/{world}: N={prob['N']} distributed one-shot {t1 - t0:.3f} s, posterior rel diff vs sequential single-GPU {err:.2e}", flush=True)
assert err < 1e-9, err
# factor left distributed (replicate=False): collective var on this rank's shard of the test points
from linpde_gp_b200 import parallel
post_d = lg.ConditionalGaussianProcess.from_observation_batches(prior, batches, nb=512, replicate=False)
lo, hi = parallel.shard_bounds(len(prob["Xt"]), rank, world)
m_s, v_s = post_d.mean(prob["Xt"][lo:hi]), post_d.var(prob["Xt"][lo:hi])
err2 = max(np.max(np.abs(m_s - m_1[lo:hi])), np.max(np.abs(v_s - v_1[lo:hi]))) / sc
c_s = post_d.cov.matrix(prob["Xt"][:40])
err3 = np.max(np.abs(c_s - post1.cov.matrix(prob["Xt"][:40]))) / sc
print(f"rank {rank}/{world}: distributed factor (not replicated): mean/var rel diff {err2:.2e}, cov rel diff {err3:.2e}", flush=True)
assert err2 < 1e-9 and err3 < 1e-9, (err2, err3)
dist.destroy_process_group()
