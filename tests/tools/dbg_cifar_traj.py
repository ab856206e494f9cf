"""debug aid: divergence of the TF32 trajectory from the fp32 one, step by step (CIFAR10CNN, SGD)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import flb200
from flb200.training import BatchedClientTrainer
from oracle import models as OM
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for steps in (1, 2, 3, 6, 12):
    outs = {}
    for prec in ("fp32", "tf32"):
        eng = BatchedClientTrainer("cifar10_cnn", 1, dev, batch_size=B, dropout_rate=0.0, precision=prec)
        eng.set_global_row(eng.layout.flatten(OM.init_weights("cifar10_cnn", 5), dev))
        g = torch.Generator().manual_seed(70)
        x = torch.randn((B * steps, 3, 32, 32), generator=g); y = torch.randint(0, 10, (B * steps,), generator=g)
        eng.load_data([x], [y])
        w0 = eng.W.clone()
        eng.train(1, 1e-2, "sgd")
        outs[prec] = eng.W.clone() - w0
    d32, dtf = outs["fp32"], outs["tf32"]
    lay = eng.layout
    per = {n: float((dtf[0, lay.offsets[n]:lay.offsets[n] + lay.shapes[n][0] * (1 if len(lay.shapes[n]) == 1 else int(torch.tensor(lay.shapes[n][1:]).prod()))] - d32[0, lay.offsets[n]:lay.offsets[n] + lay.shapes[n][0] * (1 if len(lay.shapes[n]) == 1 else int(torch.tensor(lay.shapes[n][1:]).prod()))]).norm() / d32[0, lay.offsets[n]:lay.offsets[n] + lay.shapes[n][0] * (1 if len(lay.shapes[n]) == 1 else int(torch.tensor(lay.shapes[n][1:]).prod()))].norm().clamp(min=1e-30)) for n in ("conv1.weight", "conv2.weight", "conv6.weight", "fc1.weight", "fc3.weight")}
    print(steps, "steps: rel L2 of update %.3f" % float((dtf - d32).norm() / d32.norm()), {k: round(v, 3) for k, v in per.items()})
