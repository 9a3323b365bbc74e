"""Where does the host time of SearchPipeline.submit go?  cProfile over 64 submits at 512 trees."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hanabizero_b200.mcts import MCTS, SearchConfig, SearchPipeline
from hanabizero_b200.model import MuZeroNetFull

dev = torch.device("cuda"); N, A, S, F = int(os.environ.get("N", "512")), 20, 50, 512
torch.manual_seed(0)
model = MuZeroNetFull(785 * 4, A).randomize_heads().to(dev).eval()
cfg = SearchConfig(num_simulations=S, amp_type="torch_amp")
pipe = SearchPipeline(MCTS(cfg), model, N, A, depth=8)
rng = np.random.default_rng(0)
noise = torch.from_numpy(rng.dirichlet([0.3] * A, N).astype(np.float32)).to(dev)
args = (0.25, noise, torch.zeros(N, device=dev), torch.randn(N, A, device=dev), torch.ones(N, A, dtype=torch.int32, device=dev),
        torch.rand(N, F, device=dev).half())
for _ in range(32):
    pipe.submit(*args)
pipe.drain()
pr = cProfile.Profile()
pr.enable()
for _ in range(64):
    pipe.submit(*args)
pr.disable()
pipe.drain()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(45)
