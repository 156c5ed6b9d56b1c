"""-m gpu: whole-network parity of the CUDA path against the CPU oracle on shared seeded weights and
identical synthetic inputs, and against the reference-generated golden vectors.

Tolerances (BASELINE.json north_star): heat maps within 2e-2 relative error in bf16, soft-argmax
coordinates within 0.05 px, argmax indices bit-exact on identical fp32 heat maps.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_TOL = 2e-2
PX_TOL = 0.05


def _model(width, variant, sharp=False, trainable=False, perturb=True):
    from oracle import fixtures
    from hrnet_b200.config import make_cfg
    from hrnet_b200.models import pose_hrnet, pose_hrnet_softmax
    cfg = make_cfg(width, softmax=(variant == "softmax"), trainable_softmax=trainable)
    torch.manual_seed(0)
    m = (pose_hrnet_softmax if variant == "softmax" else pose_hrnet).get_pose_net(cfg, is_train=False)
    sd = m.state_dict()
    if perturb:
        fixtures.perturb_state_dict(sd)
    if sharp:
        fixtures.sharpen_head(sd)
    if variant == "softmax" and perturb:
        sd["trainable_temp"].fill_(1.7)
    m.load_state_dict(sd)
    return m.eval(), cfg, sd


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _report(name, **kv):
    """append measured parity numbers to gpurun_out/parity_report.jsonl (copied into profiles/ by hand)"""
    import json
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
        f.write(json.dumps(dict(case=name, **kv)) + "\n")


# SPEC set   : the north_star's case - shared random-init (reference default init, seed 0) weights.
#              heat maps <= 2e-2 relative, soft-argmax <= 0.05 px.
# STRESS sets: BN statistics/affine randomised (+ temperature 1.7, + head x50 for "sharp").  The reference itself
#              under torch bf16 autocast is at 1.0e-2 (heat) / 2.1e-2 (logits) / 0.07 px on the perturbed set
#              [measured with oracle/ on CPU], i.e. at the edge of the spec numbers, so these sets use documented
#              looser bounds; they exist to catch layout/fold/indexing bugs, which produce O(1) errors.
CASES = [
    # name, width, sharp, perturb, H, W, B, heat_tol, logit_tol, px_tol
    ("hrnet_w32_softmax_default", 32, False, False, 256, 256, 1, REL_TOL, REL_TOL, PX_TOL),
    ("hrnet_w32_softmax", 32, False, True, 256, 256, 1, 5e-2, 4e-2, 0.15),
    ("hrnet_w32_softmax_sharp", 32, True, True, 256, 256, 1, None, 4e-2, None),
    ("hrnet_w48_softmax_rect", 48, False, True, 128, 96, 2, 5e-2, 4e-2, 0.15),
]


@pytest.mark.parametrize("name,width,sharp,perturb,H,W,B,heat_tol,logit_tol,px_tol", CASES, ids=[c[0] for c in CASES])
def test_softmax_variant_matches_oracle_and_golden(golden_dir, name, width, sharp, perturb, H, W, B, heat_tol,
                                                   logit_tol, px_tol):
    from oracle import decode_oracle, fixtures, hrnet_oracle
    from hrnet_b200.utils.heatmap_decoding import get_final_preds
    from hrnet_b200.core.inference import get_max_preds
    m, cfg, sd = _model(width, "softmax", sharp, perturb=perturb)
    x = fixtures.images(B, H, W)
    arch = hrnet_oracle.Arch.from_cfg(cfg)
    o_heat, o_feat, o_temp, o_logits = hrnet_oracle.forward(sd, x, arch, "softmax")
    m = m.cuda()
    heat, feat, temp = m(x.cuda())
    logits = m.engine().plan(B, H, W).out["logits"].clone()
    torch.cuda.synchronize()
    assert heat.shape == o_heat.shape and feat.shape == o_feat.shape
    assert float(temp) == pytest.approx(1.7 if perturb else 1.0)
    coords = get_final_preds(heat, True).cpu().numpy()
    o_coords = decode_oracle.spatial_expectation2d(o_heat.numpy())
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    r = dict(feat=_rel(feat.cpu(), o_feat), logits=_rel(logits.cpu(), o_logits), heat=_rel(heat.cpu(), o_heat),
             px=float(np.abs(coords - o_coords).max()), px_vs_reference=float(np.abs(coords - g["soft_coords"]).max()),
             heat_vs_reference=_rel(heat.cpu()[:, :, ::2, ::2], torch.from_numpy(g["heat"])))
    _report(name, **r)
    assert r["feat"] < 4e-2 and r["logits"] < logit_tol, r
    if heat_tol is not None:
        assert r["heat"] < heat_tol and r["heat_vs_reference"] < heat_tol, r
    if px_tol is not None:
        assert r["px"] < px_tol and r["px_vs_reference"] < px_tol, r
    # the fused decode inside the plan agrees with the stand-alone decode
    assert np.abs(m.engine().plan(B, H, W).out["coords"].cpu().numpy() - coords).max() < 1e-3
    # argmax is bit-exact on identical fp32 heat maps (feed the ORACLE maps to the CUDA decoder)
    p, mv = get_max_preds(o_heat.numpy())
    op, omv = decode_oracle.get_max_preds(o_heat.numpy())
    assert np.array_equal(p, op) and np.array_equal(mv, omv)
    if sharp:   # peaky maps: end-to-end argmax lands on the oracle's pixel (or a neighbour) almost everywhere
        p2, _ = get_max_preds(heat.cpu().numpy())
        frac = float((np.abs(p2 - op).max(-1) <= 1).mean())
        _report(name + "_argmax", within_1px=frac, exact=float((p2 == op).all(-1).mean()))
        assert frac > 0.75      # x50 head: near-tied peaks flip under bf16 logit noise (the reference under bf16 autocast does too)


def test_raw_variant_matches_oracle_and_golden(golden_dir):
    from oracle import fixtures, hrnet_oracle
    m, cfg, sd = _model(32, "raw")
    x = fixtures.images(1)
    arch = hrnet_oracle.Arch.from_cfg(cfg)
    o_logits, o_feat = hrnet_oracle.forward(sd, x, arch, "raw")
    m = m.cuda()
    logits, feat = m(x.cuda())
    torch.cuda.synchronize()
    assert logits.shape == (1, 21, 64, 64) and feat.shape == (1, 32, 64, 64)
    assert _rel(logits.cpu(), o_logits) < REL_TOL
    assert _rel(feat.cpu(), o_feat) < REL_TOL
    g = np.load(os.path.join(golden_dir, "hrnet_w32_raw.npz"))
    assert _rel(logits.cpu()[:, :, ::2, ::2], torch.from_numpy(g["logits"])) < REL_TOL


def test_batch_rows_are_independent_and_deterministic():
    """Images are independent units: row b of a batch equals the same image run alone; replay is bit-stable."""
    from oracle import fixtures
    m, _, _ = _model(32, "softmax")
    m = m.cuda()
    x = fixtures.images(3).cuda()
    h3 = m(x)[0].clone()
    h3b = m(x)[0].clone()
    assert torch.equal(h3, h3b)
    # a different batch size may pick different tile shapes (K-chunking changes the fp32 summation order), so
    # rows agree to bf16 noise rather than bit for bit
    h1 = m(x[1:2])[0].clone()
    assert _rel(h3[1:2].cpu(), h1.cpu()) < 5e-2


def test_graph_and_eager_paths_agree():
    from oracle import fixtures
    m, _, _ = _model(32, "softmax")
    m = m.cuda()
    x = fixtures.images(2).cuda()
    a = m(x)[0].clone()
    m.engine().use_graph = False
    b = m(x)[0].clone()
    assert torch.equal(a, b)


def test_streaming_predictor_matches_direct_calls():
    from oracle import fixtures
    from hrnet_b200.pipeline import StreamingPredictor
    from hrnet_b200.utils.heatmap_decoding import get_final_preds
    m, _, _ = _model(32, "softmax")
    m = m.cuda()
    batches = [fixtures.images(2, 128, 128, seed=20 + i).pin_memory() for i in range(5)]
    direct = [get_final_preds(m(b.cuda())[0], True).cpu().clone() for b in batches]
    outs = [o.clone() for o in StreamingPredictor(m).run(iter(batches))]
    assert len(outs) == 5
    for a, b in zip(outs, direct):
        assert torch.equal(a, b)


def test_data_parallel_replicas_share_one_engine_per_device():
    """nn.DataParallel (the reference's default wrapper, tools/train.py:250-254) replicates the module on every forward; the
    replicas must find the per-device engine of the ORIGINAL module instead of re-packing the weights on every call, must
    read their (non-leaf, broadcast) parameters, and must refuse to train.  (Replica logic on one device, called in turn;
    the concurrent one-thread-per-GPU form runs in tests/test_gpu_multi.py on two devices.)"""
    from torch.nn.parallel import replicate
    from oracle import fixtures
    m, _, _ = _model(32, "softmax")
    m = m.cuda()
    x = fixtures.images(2, 128, 128).cuda()
    ref = m(x)[0].clone()
    built = []
    for _ in range(2):
        reps = replicate(m, [0, 0])
        for r in reps:
            assert len(list(r.parameters())) == 0            # what torch does to replicas: the engine must not rely on it
            assert torch.equal(r(x)[0], ref)
        built.append(id(m._shared["engines"][torch.device("cuda", 0)][1]))
    assert built[0] == built[1]
    with torch.no_grad():
        m.last_layer[3].bias.add_(1.0)                       # parameter change on the original -> replicas re-pack
    reps = replicate(m, [0])
    assert not torch.equal(reps[0](x)[0], ref)
    assert id(m._shared["engines"][torch.device("cuda", 0)][1]) != built[0]
    with pytest.raises(RuntimeError, match="DataParallel"):
        replicate(m.train(), [0])[0](x)
