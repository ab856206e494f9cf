"""debug aid: graph-replayed round time vs clients per GPU (fixed cost of the launch sequence vs marginal cost per client)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flb200
from flb200.models_pytorch import ModelFactory
from flb200.simulation import FederatedRoundEngine
dev = torch.device("cuda:0")
model = sys.argv[1] if len(sys.argv) > 1 else "simple_cnn"
for K in [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,2,5,10,20,40").split(",")]:
    eng = FederatedRoundEngine(model, K, dev, dp_mode="update", precision="tf32")
    torch.manual_seed(0)
    eng.set_global_weights(ModelFactory.create_model(model).get_model_weights())
    eng.load_synthetic()
    for _ in range(4):
        eng.run_round(read_metrics=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.run_round(read_metrics=False)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    steps = eng.trainer.max_steps()
    print(f"K={K:3d}  round {ms:7.3f} ms  ({steps} steps, {ms / steps * 1e3:6.1f} us/step)  {eng.samples_per_round() / ms * 1e3:10.0f} samples/s")
    del eng
    torch.cuda.empty_cache()
