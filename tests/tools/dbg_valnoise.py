"""Run-to-run spread of the LocalTrainer trajectory check of tests/test_gpu_round2.py::test_validation_loader_leaves_optimizer_state_alone
(two identical 3-epoch Adam runs, one with a validation loader): relative L2 of the difference over the accumulated update."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import flb200  # noqa
from oracle import models as OM, training as OT
from flb200.models_pytorch import ModelFactory
from flb200.training import LocalTrainer
MODEL = "simple_cnn"
dev = torch.device("cuda:0")
def _data(seed, n):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((n,) + OM.input_shape(MODEL), generator=g), torch.randint(0, 10, (n,), generator=g)
x, y = _data(22, 40); xv, yv = _data(23, 24)
batches = OT.make_batches(x, y, 8); val = OT.make_batches(xv, yv, 8)
w0 = OM.init_weights(MODEL, 12)
vals = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    out = []
    for vl in ((None, None) if os.environ.get("DBG_NOVAL") else (None, val)):
        model = ModelFactory.create_model(MODEL, dropout_rate=0.0)
        model.set_model_weights(OM.init_weights(MODEL, 12))
        LocalTrainer(model, dev).train_local_model(batches, 3, learning_rate=1e-3, optimizer_type="adam", validation_loader=vl, save_checkpoints=False)
        out.append({k: v.cpu() for k, v in model.get_model_weights().items()})
    num = sum(float(((out[1][n] - out[0][n]).double() ** 2).sum()) for n in w0)
    den = sum(float(((out[0][n] - w0[n]).double() ** 2).sum()) for n in w0)
    vals.append((num / den) ** 0.5)
print(os.environ.get("FLB_NO_L2_PERSIST"), ["%.2e" % v for v in vals])
