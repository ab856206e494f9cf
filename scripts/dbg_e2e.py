"""debug aid: end-to-end round time for different orderings of upload / round / readback (see bench.py e2e loop)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flb200
from flb200.models_pytorch import ModelFactory
from flb200.simulation import FederatedRoundEngine, synthetic_client_data, synthetic_num_samples
dev = torch.device("cuda:0")
K = 10
eng = FederatedRoundEngine("simple_cnn", K, dev, dp_mode="update", precision="tf32")
torch.manual_seed(0)
eng.set_global_weights(ModelFactory.create_model("simple_cnn").get_model_weights())
host = [synthetic_client_data("simple_cnn", i) for i in range(K)]
sizes = [synthetic_num_samples("simple_cnn", i) for i in range(K)]
x_host = torch.cat([h[0].reshape(h[0].shape[0], -1) for h in host]).pin_memory()
y_host = torch.cat([h[1] for h in host]).to(torch.int32).pin_memory()
gw = torch.empty(eng.layout.P, dtype=torch.float32).pin_memory()
eng.load_packed(x_host, y_host, sizes)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run(mode, n=20):
    tot = 0.0
    for i in range(n + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "device":
            eng.run_round(read_metrics=False)
        elif mode == "old":
            eng.use_prefetched(); eng.prefetch_packed(x_host, y_host)
            eng.run_round(read_metrics=True)
            gw.copy_(eng.global_row[:eng.layout.P], non_blocking=True)
        elif mode == "new":
            eng.use_prefetched(); eng.start_round(); eng.prefetch_packed(x_host, y_host)
            eng.finish_round(True, gw)
        elif mode == "new_prefetch_first":
            eng.use_prefetched(); eng.prefetch_packed(x_host, y_host); eng.start_round()
            eng.finish_round(True, gw)
        elif mode == "no_upload":
            eng.start_round(); eng.finish_round(True, gw)
        elif mode == "no_readback":
            eng.use_prefetched(); eng.start_round(); eng.prefetch_packed(x_host, y_host)
            eng.finish_round(False, None)
        e1.record(); e1.synchronize()
        if i >= 3:
            tot += e0.elapsed_time(e1)
    return tot / n

eng.prefetch_packed(x_host, y_host)
for m in ["device", "old", "new", "new_prefetch_first", "no_upload", "no_readback", "old", "new"]:
    print(f"{m:20s} {run(m):.4f} ms")
