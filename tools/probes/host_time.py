"""How much of a PPO update is host time?  cProfile around PPO.update on the C2 shape: everything that is not the final device->host
read (`.tolist()`, where the host waits for the device) is time the Python side spent enqueuing the 12 500 launches."""
import cProfile, pstats, io, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ppodash_b200 as ppd


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


dev = "cuda:0"
T, N = 512, 32
torch.manual_seed(0)
pol = ppd.Policy((3, 84, 84), Discrete(8), base_kwargs={"recurrent": True}, vector_obs_len=15).to(dev)
pol.engine("tf32x3")
st = ppd.RolloutStorage(T, N, (3, 84, 84), [15], Discrete(8), 512)
st.to(dev)
st.obs.normal_(); st.vector_obs.normal_(); st.rewards.normal_(); st.value_preds.normal_(); st.action_log_probs.fill_(-2.0)
st.actions.random_(0, 8); st.masks.fill_(1.0); st.bad_masks.fill_(1.0)
agent = ppd.algo.PPO(pol, 0.1, 8, 8, 0.5, 0.001, lr=1e-4, eps=1e-5, max_grad_norm=0.5)
nv = torch.zeros(N, 1, device=dev)
for _ in range(2):
    st.compute_returns(nv, True, 0.99, 0.95, False)
    agent.update(st)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
st.compute_returns(nv, True, 0.99, 0.95, False)
agent.update(st)
pr.disable()
wall = time.perf_counter() - t0
s = io.StringIO()
ps = pstats.Stats(pr, stream=s).sort_stats("tottime")
ps.print_stats(14)
txt = s.getvalue()
wait = 0.0
for line in txt.splitlines():
    if "tolist" in line:
        wait = float(line.split()[1])
print(f"wall {1e3 * wall:.1f} ms (under cProfile), of which waiting in .tolist() {1e3 * wait:.1f} ms -> host enqueue time {1e3 * (wall - wait):.1f} ms")
print("\n".join(txt.splitlines()[:30]))
