"""debug aid: per-kernel step breakdown (CUDA events, eager) for different client counts -> fixed vs per-tile cost"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flb200
from flb200 import _lib as L
from flb200.training import BatchedClientTrainer
from flb200.models_pytorch import ModelFactory
dev = torch.device("cuda:0")
model = sys.argv[1] if len(sys.argv) > 1 else "simple_cnn"
for K in [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,10,40").split(",")]:
    eng = BatchedClientTrainer(model, K, dev, batch_size=32, precision="tf32")
    torch.manual_seed(0)
    eng.set_global_row(eng.layout.flatten(ModelFactory.create_model(model).get_model_weights(), dev))
    shp = (1, 28, 28) if model == "simple_cnn" else (3, 32, 32)
    eng.load_data([torch.randn((96,) + shp) for _ in range(K)], [torch.randint(0, 10, (96,)) for _ in range(K)])
    eng._fill_args(1e-3, "adam")
    L.call("flb_train_begin_epoch", C.byref(eng.args), L.stream_ptr(dev))
    eng.profile_step()
    L.call("flb_train_begin_epoch", C.byref(eng.args), L.stream_ptr(dev))
    acc = {}
    for _ in range(2):
        for k, v in eng.profile_step().items():
            acc[k] = acc.get(k, 0) + v / 2
    print(K, {k: round(v * 1e3, 1) for k, v in acc.items()}, "sum_us", round(sum(acc.values()) * 1e3))
    del eng
    torch.cuda.empty_cache()
