import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import flb200
from flb200.training import BatchedClientTrainer
from oracle import models as OM, training as OT
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(22)
x = torch.randn((40,1,28,28), generator=g); y = torch.randint(0,10,(40,),generator=g)
w0 = OM.init_weights("simple_cnn", 12)
wref = {k:v.clone() for k,v in w0.items()}
OT.train_local_model("simple_cnn", wref, OT.make_batches(x,y,8), 2, 1e-3, "adam")
def run(graph):
    eng = BatchedClientTrainer("simple_cnn", 1, dev, batch_size=8, dropout_rate=0.0, precision="fp32", use_graph=graph)
    eng.set_client_weights(0, w0); eng.load_data([x],[y]); eng.train(2, 1e-3, "adam")
    return eng.client_weights(0, "cpu")
a = run(False); b = run(False); c = run(True)
for k in w0:
    print(k, "eager-vs-ref %.2e  eager-vs-eager %.2e  graph-vs-ref %.2e" % ((a[k]-wref[k]).abs().max(), (a[k]-b[k]).abs().max(), (c[k]-wref[k]).abs().max()))
