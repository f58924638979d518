"""Dev tool (GPU): where the start-up of a batch goes -- eager prefill and per-batch-size CUDA-graph capture."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from vox_serve_b200.model.orpheus import OrpheusModel  # noqa: E402
from vox_serve_b200.requests import Request  # noqa: E402
from vox_serve_b200.scheduler import Scheduler  # noqa: E402
from vox_serve_b200.worker import ModelWorker  # noqa: E402

model = OrpheusModel("orpheus-synthetic:0", device="cuda:0", mask_stop_token=True, max_tokens=1200)
worker = ModelWorker("orpheus-synthetic", max_batch_size=32, max_num_pages=2048, page_size=128, model=model,
                     max_prefill_tokens=1024)
sched = Scheduler(worker)
g = torch.Generator().manual_seed(1)
for i in range(6):
    sched.submit(Request(request_id=f"r{i}", prompt=torch.randint(0, 128000, (128,), generator=g).tolist(),
                         model_kwargs={"voice": None}))
pr = cProfile.Profile()
for step in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if step == 6:
        pr.enable()
    sched._step()
    if step == 6:
        pr.disable()
    torch.cuda.synchronize()
    print(f"step {step}: {(time.perf_counter() - t0) * 1e3:.1f} ms, graphs {sorted(worker.decode_graphs)}")
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
