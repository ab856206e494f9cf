"""GPU parity of the update codecs behind ModelCompressionService (src/shared/compression.py): uint8 affine quantisation
and exact top-k sparsification, against the numpy oracle and the golden vectors written by the unmodified reference."""
import pickle

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import compression as OC
from oracle import models as OM

pytestmark = pytest.mark.gpu


def _select(x_rows, seg, counts, dev):
    from flb200 import ops
    x = torch.from_numpy(x_rows).to(dev)
    ld = (x.shape[1] + 31) // 32 * 32
    rows = torch.zeros((x.shape[0], ld), dtype=torch.float32, device=dev)
    rows[:, :x.shape[1]] = x
    seg_t = torch.tensor(seg, dtype=torch.int64, device=dev)
    idx, val, off_t, kk_t = ops.topk_select(rows, seg_t, counts, P=x.shape[1])
    return rows, seg_t, idx, val, off_t, kk_t


@pytest.mark.parametrize("sp", [0.9, 0.5, 0.99995])
def test_topk_matches_reference_golden(cuda_device, sp):
    from flb200 import ops
    gold = load_golden("codecs.npz")
    x = gold["x"]
    k = OC.topk_count(x.size, sp)
    rows, seg_t, idx, val, off_t, kk_t = _select(x[None, :], [0, x.size], [k], cuda_device)
    got_idx = idx[0, :k].cpu().numpy()
    assert k == len(gold[f"topk{sp}/idx"])
    assert set(got_idx.tolist()) == set(gold[f"topk{sp}/idx"].tolist())
    assert np.all(np.diff(got_idx) > 0)                       # index order
    np.testing.assert_array_equal(val[0, :k].cpu().numpy(), x[got_idx])
    dense = ops.topk_scatter(idx, val, seg_t, kk_t, off_t, x.size)
    np.testing.assert_array_equal(dense[0, :x.size].cpu().numpy(), gold[f"topk{sp}/dense"])


def test_topk_batched_layers_clients_and_ties(cuda_device):
    """3 clients x 4 ragged layers in ONE launch; heavy ties (quantised magnitudes, +-0, a constant layer)."""
    rng = np.random.default_rng(5)
    sizes = [1, 37, 5000, 70001]
    seg = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    K = 3
    x = np.zeros((K, seg[-1]), dtype=np.float32)
    for c in range(K):
        x[c, seg[0]:seg[1]] = rng.standard_normal(1)
        x[c, seg[1]:seg[2]] = 0.5                                        # constant: pure tie-break by index
        x[c, seg[2]:seg[3]] = np.round(rng.standard_normal(5000) * 4) / 4 * rng.choice([-1, 1], 5000)
        x[c, seg[3]:seg[4]] = rng.standard_normal(70001) * (rng.random(70001) < 0.7)   # 30 % exact zeros (+0/-0)
    counts = [1, 10, 1234, 7000]
    rows, seg_t, idx, val, off_t, kk_t = _select(x, seg, counts, cuda_device)
    offs = np.concatenate([[0], np.cumsum(counts)])
    for c in range(K):
        for l in range(4):
            seg_x = x[c, seg[l]:seg[l + 1]]
            _, ref_idx = OC.sparsify(seg_x, 0.0)                          # full stable order by -|x|
            ref = np.sort(ref_idx[:counts[l]])
            got = idx[c, offs[l]:offs[l + 1]].cpu().numpy()
            np.testing.assert_array_equal(got, ref, err_msg=f"client {c} layer {l}")
            np.testing.assert_array_equal(val[c, offs[l]:offs[l + 1]].cpu().numpy(), seg_x[ref])


@pytest.mark.parametrize("algo,kw", [("quantization", {"bits": 8}), ("quantization", {"bits": 4, "symmetric": False}),
                                     ("topk", {"sparsity_ratio": 0.9})])
def test_compression_service_round_trip(cuda_device, algo, kw):
    from flb200.compression import CompressionError, ModelCompressionService, create_compression_service
    w = {k: v.to(cuda_device) for k, v in OM.init_weights("simple_cnn", 4).items()}
    svc = create_compression_service(algo, **kw)
    blob = svc.compress_weights(w)
    pkg = pickle.loads(blob)
    assert set(pkg) == {"compressed_data", "metadata"} and pkg["metadata"]["algorithm"] == svc.compressor.get_compression_name()
    out = ModelCompressionService("topk" if algo == "quantization" else "quantization").decompress_weights(blob)   # picks the codec from metadata
    assert list(out) == list(w) and all(v.is_cuda and v.shape == w[k].shape for k, v in out.items())
    for name, t in w.items():
        x = t.cpu().numpy().reshape(-1)
        if algo == "quantization":
            q, s, z = OC.quantize(x, kw.get("bits", 8), kw.get("symmetric", True))
            ref = OC.dequantize(q, s, z)
            meta = pkg["metadata"]["quantization_params"][name]
            assert abs(meta["scale"] - s) <= 1e-6 * abs(s) and meta["zero_point"] == z
            np.testing.assert_allclose(out[name].cpu().numpy().reshape(-1), ref, rtol=0, atol=1e-6 * max(abs(s), 1e-12) + 1e-9)
        else:
            vals, idx = OC.sparsify(x, kw["sparsity_ratio"])
            np.testing.assert_array_equal(out[name].cpu().numpy().reshape(-1), OC.desparsify(vals, idx, x.shape))
    assert 0.0 < svc.estimate_compression_ratio(w) < (0.5 if algo == "quantization" else 0.6)
    with pytest.raises(CompressionError):
        svc.compress_weights({k: v.cpu() for k, v in w.items()})          # no CPU fallback
    with pytest.raises(ValueError, match="Unknown compression algorithm"):
        ModelCompressionService("zstd")


def test_q8_scale_from_the_dp_pass_is_bit_identical(cuda_device):
    """f2: the clip + noise pass reduces the quantiser's per-(client, layer) max|upload| on the side
    (flb_dp_clip_noise_absmax); quantising from it must give exactly the codes / scales of the stand-alone quantiser run on
    the same upload rows -- including quads that straddle a layer boundary, an all-zero layer and NaN-free ragged tails."""
    from flb200 import ops
    g = torch.Generator().manual_seed(3)
    offs = [0, 5, 5 + 288, 5 + 288 + 33, 5 + 288 + 33 + 1031, 5 + 288 + 33 + 1031 + 4096, 5 + 288 + 33 + 1031 + 4096 + 7]   # unaligned boundaries
    P, K = offs[-1], 3
    ld = (P + 31) // 32 * 32
    local = torch.zeros((K, ld))
    local[:, :P] = torch.randn((K, P), generator=g) * 0.05
    glob = torch.zeros(ld)
    glob[:P] = torch.randn(P, generator=g) * 0.05
    local[:, offs[2]:offs[3]] = glob[offs[2]:offs[3]]              # layer 2: zero delta and zero global -> all-zero layer below
    glob[offs[2]:offs[3]] = 0
    local[:, offs[2]:offs[3]] = 0
    local, glob = local.to(cuda_device), glob.to(cuda_device)
    seg = torch.tensor(offs, dtype=torch.int64, device=cuda_device)
    z = torch.zeros((K, ld), device=cuda_device)                     # sigma * 0: the upload is global + clipped delta
    up_a, norms_a = ops.dp_clip_noise(local, glob, 0.5, 4.8448, seed=1, z=z, P=P)
    up_b, norms_b, absmax = ops.dp_clip_noise(local, glob, 0.5, 4.8448, seed=1, z=z, P=P, absmax_seg=seg)
    assert torch.equal(up_a[:, :P], up_b[:, :P]) and torch.equal(norms_a, norms_b)      # columns past P are never written
    want = torch.stack([up_a[:, offs[i]:offs[i + 1]].abs().max(dim=1).values for i in range(len(offs) - 1)], dim=1)
    assert torch.equal(absmax.view(torch.float32), want)
    qa, sa, za = ops.q8_quantize(up_a, seg, P=P)
    qb, sb, zb = ops.q8_quantize(up_b, seg, P=P, absmax=absmax)
    assert torch.equal(sa, sb) and torch.equal(za, zb)
    assert torch.equal(qa[:, :P], qb[:, :P])
    # Philox noise path: same statement with generated noise (the absmax sees exactly the values that were written)
    up_c, _, absmax_c = ops.dp_clip_noise(local, glob, 0.5, 4.8448, seed=9, stream_base=5, P=P, absmax_seg=seg)
    want_c = torch.stack([up_c[:, offs[i]:offs[i + 1]].abs().max(dim=1).values for i in range(len(offs) - 1)], dim=1)
    assert torch.equal(absmax_c.view(torch.float32), want_c)
