"""GPU: the model graphs on the CUDA path against (1) golden outputs of the imported reference
(CPU-eager, tests/golden/model_*.npz) and (2) the oracle model + oracle training iteration run
GPU-eager on the same box.  Quantisation is discrete: a weight/activation code that sits on a
rounding tie may flip between the two paths (cuDNN vs CPU conv summation order, 1-ulp stats), so the
model-level bars are norm-relative, stated per assert."""
import json
import os

import numpy as np
import pytest
import torch

import alignq_b200 as aq
from alignq_b200.model import dann, densenet, mobilenetV2, resnet
from alignq_b200.utils.train import QATStep
from oracle import models_oracle as MO

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"

BUILD = {
    "resnet20_A": ("A", lambda: resnet.resnet20_quant(8, 8, "second")),
    "resnet20_B": ("B", lambda: resnet.resnet20_quant(8, 8, "second")),
    "resnet56_B": ("B", lambda: resnet.resnet56_quant(8, 8, "second")),
    "mobilenetv2_A": ("A", lambda: mobilenetV2.mobile_v2(4, 4, "second")),
    "densenet40_A": ("A", lambda: densenet.densenet_40_quant(8, 8, "second")),
    "resnet50dann_C": ("C", lambda: dann.resnet50_dann(8, 8, "second")),
}


def relnorm(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture(autouse=True)
def _deterministic():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("job", list(BUILD))
def test_model_forward_backward_vs_reference_golden(job):
    variant, ctor = BUILD[job]
    g = np.load(os.path.join(GOLDEN, f"model_{job}.npz"))
    x, tgt = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["target"]).to(DEV)
    bits = 4 if "mobilenet" in job else 8
    aq.set_args(variant=variant, train_batch_size=x.shape[0], bitW=bits, abitW=bits, act_range=2, method="ours")
    model = ctor()
    model.load_state_dict(MO.deterministic_fill(model.state_dict(), seed=11))
    model.to(DEV).train()
    out = model(x, 0.5) if "dann" in job else model(x)
    logits = out[0] if isinstance(out, tuple) else out
    loss = torch.nn.functional.cross_entropy(logits, tgt)
    tl = out[-1] if isinstance(out, tuple) else None
    (loss if tl is None else loss + tl).backward()
    # Bars: 5x the reference's own sensitivity band (its outputs under a 1e-7 relative weight
    # perturbation, recorded by oracle/make_model_golden.py), floor 5e-3 -- quantisation is
    # discontinuous, see the module docstring.  The tight model-level check is the 32-bit wiring test.
    e = relnorm(logits.detach().cpu(), torch.from_numpy(g["logits"]))
    band = float(g["band_logits"])
    print(f"{job}: logits rel-norm err {e:.2e} (reference band {band:.2e})")
    assert e <= max(5 * band, 5e-3), "logits vs reference"
    if tl is not None:
        et = abs(float(tl.detach()) - float(g["trans_loss"])) / abs(float(g["trans_loss"]))
        print(f"{job}: trans_loss rel err {et:.2e} (reference band {float(g['band_trans_loss']):.2e})")
        assert et <= max(5 * float(g["band_trans_loss"]), 1e-3)
    fp = torch.from_numpy(g["grad_fp"])
    mine = torch.stack([p.grad.double().abs().sum() if p.grad is not None else torch.zeros((), dtype=torch.float64, device=DEV)
                        for _, p in model.named_parameters()]).cpu()
    eg = relnorm(mine, fp[:, 1])
    print(f"{job}: per-parameter |grad| sums rel-norm err {eg:.2e} (reference band {float(g['band_grad_fp']):.2e})")
    assert eg <= max(5 * float(g["band_grad_fp"]), 2e-2), "gradient magnitudes vs reference"


@pytest.mark.parametrize("job", list(BUILD))
def test_model_wiring_at_32_bit_vs_reference_golden(job):
    """bitW = abitW = 32 turns every quantizer into the identity (QA:64-67,92-95): the graph is then a
    smooth function and must match the reference's 32-bit run tightly (conv/BN/ReLU/shortcut order)."""
    variant, _ = BUILD[job]
    g = np.load(os.path.join(GOLDEN, f"model_{job}.npz"))
    x, tgt = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["target"]).to(DEV)
    aq.set_args(variant=variant, train_batch_size=x.shape[0], bitW=32, abitW=32, act_range=2, method="ours")
    ctor32 = {"resnet20": lambda: resnet.resnet20_quant(32, 32, "second"), "resnet56": lambda: resnet.resnet56_quant(32, 32, "second"),
              "mobilenetv2": lambda: mobilenetV2.mobile_v2(32, 32, "second"), "densenet40": lambda: densenet.densenet_40_quant(32, 32, "second"),
              "resnet50dann": lambda: dann.resnet50_dann(32, 32, "second")}[job.split("_")[0]]
    model = ctor32()
    model.load_state_dict(MO.deterministic_fill(model.state_dict(), seed=11))
    model.to(DEV).train()
    out = model(x, 0.5) if "dann" in job else model(x)
    logits = out[0] if isinstance(out, tuple) else out
    torch.nn.functional.cross_entropy(logits, tgt).backward()
    e = relnorm(logits.detach().cpu(), torch.from_numpy(g["logits_fp32"]))
    fp = torch.from_numpy(g["grad_fp_fp32"])
    mine = torch.stack([p.grad.double().abs().sum() if p.grad is not None else torch.zeros((), dtype=torch.float64, device=DEV)
                        for _, p in model.named_parameters()]).cpu()
    eg = relnorm(mine, fp[:, 1])
    print(f"{job} @32 bit: logits rel-norm err {e:.2e}, |grad| sums rel-norm err {eg:.2e}")
    assert e <= 1e-3 and eg <= 1e-2


@pytest.mark.parametrize("variant", ["A", "B"])
def test_training_iterations_vs_oracle_trainer_on_gpu(variant):
    """Product QATStep (CUDA kernels, multi-tensor SGD, ADMM_OPT) vs the oracle's restatement of the
    reference train() body, both on the GPU with the same cuDNN convs: 3 iterations."""
    B = 16
    aq.set_args(variant=variant, train_batch_size=B, bitW=8, abitW=8, act_range=2, method="ours", lam=1.0, lam2=4.0)
    torch.manual_seed(0)
    x = torch.randn(B, 3, 32, 32, device=DEV)
    t = torch.randint(0, 10, (B,), device=DEV)
    prod = resnet.resnet20_quant(8, 8, "second")
    sd = MO.deterministic_fill(prod.state_dict(), seed=3)
    prod.load_state_dict(sd)
    prod.to(DEV).train()
    orc = MO.OracleResNet([3, 3, 3], 8, 8, variant, 2.0, dim=B)
    orc.load_state_dict(sd)
    orc.to(DEV).train()
    # live sensitivity band: the same oracle loop started from weights perturbed by a few 1e-7 relative.  One
    # perturbed run is a noisy estimate (whether a rounding tie flips is all-or-nothing), so take the worst of four.
    perts = []
    for scale in (1e-7, -1e-7, 3e-7, -3e-7):
        pm = MO.OracleResNet([3, 3, 3], 8, 8, variant, 2.0, dim=B)
        pm.load_state_dict(sd)
        pm.to(DEV).train()
        with torch.no_grad():
            for p in pm.parameters():
                p.mul_(1.0 + scale)
        perts.append(pm)
    step = QATStep(prod, lr=0.04, momentum=0.9, weight_decay=1e-4)
    tr = MO.OracleTrainer(orc, lr=0.04, momentum=0.9, weight_decay=1e-4, lam=1.0, lam2=4.0, bitW=8)
    trps = [MO.OracleTrainer(pm, lr=0.04, momentum=0.9, weight_decay=1e-4, lam=1.0, lam2=4.0, bitW=8) for pm in perts]

    def worst(a, b):
        return max(relnorm(p.detach(), q.detach()) for p, q in zip(a.parameters(), b.parameters()))

    for it in range(3):
        lp = float(step.step(x, t))
        lo = float(tr.step(x, t)[0])
        lbs = [float(trp.step(x, t)[0]) for trp in trps]
        err, band = worst(prod, orc), max(worst(pm, orc) for pm in perts)
        lband = max(abs(lb - lo) for lb in lbs)
        print(f"variant {variant} it {it}: CE {lp:.6f} vs oracle {lo:.6f} (perturbed oracles within {lband:.2e}); "
              f"param err {err:.2e} (band {band:.2e})")
        if it == 0:
            assert abs(lp - lo) <= 1e-4 * abs(lo), "first-iteration CE loss"
            # p.grad left behind by SGD.step: the surrogate for quantized convs (optimizer.py:232-249)
            assert relnorm(prod.layers[0].conv0.weight.grad, orc.layers[0].conv0.weight.grad) <= max(band, 1e-3)
            assert relnorm(prod.logit.weight.grad, orc.logit.weight.grad) <= max(band, 1e-3)
        assert err <= max(3 * band, 1e-4), f"iteration {it}: product drifts from the oracle faster than a 1-ulp perturbation"
        assert abs(lp - lo) <= max(3 * lband, 1e-4 * abs(lo)), f"iteration {it}: CE loss outside the band"

def test_graph_replay_equals_eager():
    B = 32
    aq.set_args(variant="A", train_batch_size=B, bitW=8, abitW=8, act_range=2)
    torch.manual_seed(1)
    x = torch.randn(B, 3, 32, 32, device=DEV)
    t = torch.randint(0, 10, (B,), device=DEV)
    models = []
    for _ in range(4):
        m = resnet.resnet20_quant(8, 8, "second")
        m.load_state_dict(MO.deterministic_fill(m.state_dict(), seed=5))
        models.append(m.to(DEV).train())
    with torch.no_grad():                  # two eager twins started 1e-7 away: the live sensitivity band
        for k, sc in ((2, 1e-7), (3, -1e-7)):
            for p in models[k].parameters():
                p.mul_(1.0 + sc)
    eager, graphed = QATStep(models[0]), QATStep(models[1])
    twins = [QATStep(models[2]), QATStep(models[3])]
    graphed.capture(x, t, warmup=3)        # 3 real warm-up iterations; the capture pass itself does not execute
    lg = graphed.step(x, t).clone()        # iteration 4 by replay (step() returns a live buffer)
    for _ in range(4):
        le = eager.step(x, t)
        lt = [tw.step(x, t) for tw in twins]
    # same kernels, same order; cuDNN may pick a different algorithm under capture (and its wgrad is not
    # deterministic), and the loop is chaotic (see test_training_iterations_vs_oracle_trainer_on_gpu): the bar is
    # what a 1e-7 perturbation of the start does to the same eager loop, with a floor
    band = max(abs(float(v) - float(le)) for v in lt)
    assert abs(float(le) - float(lg)) <= max(5 * band, 5e-2 * abs(float(le)))
    assert torch.isfinite(lg)
    l5 = float(graphed.step(x, t))
    assert l5 == l5 and l5 != float(lg)    # replay advances the optimisation (state lives outside the graph)


def test_weight_bank_matches_per_layer_quantization():
    """One multi-tensor launch for all conv weights == 21 per-layer launches (same kernels; the bank's
    slices are not all 16-byte aligned, so a chunk may take the scalar instead of the 128-bit path and
    sum its fp64 partials in another order: results agree to fp32 round-off, not always bit for bit)."""
    B = 16
    aq.set_args(variant="A", train_batch_size=B, bitW=8, abitW=8, act_range=2)
    torch.manual_seed(4)
    x = torch.randn(B, 3, 32, 32, device=DEV)
    t = torch.randint(0, 10, (B,), device=DEV)
    losses, params = [], []
    for bank in (False, True):
        m = resnet.resnet20_quant(8, 8, "second")
        m.load_state_dict(MO.deterministic_fill(m.state_dict(), seed=9))
        m.to(DEV).train()
        st = QATStep(m, bank_weights=bank)
        assert (st.bank is not None) == bank
        losses.append([float(st.step(x, t))])
        params.append([p.detach().clone() for p in m.parameters()])        # after ONE update (the loop is chaotic)
        losses[-1] += [float(st.step(x, t)) for _ in range(2)]
        if bank:
            assert list(m.state_dict().keys())[0] == "conv0.weight" and m.conv0.weight.shape == (16, 3, 3, 3)
            assert m.layers[0].conv0.quantize_fn.weight_pdf.shape == m.layers[0].conv0.weight.shape
    assert abs(losses[0][0] - losses[1][0]) <= 1e-6 * abs(losses[0][0])
    assert all(abs(a - b) <= 2e-2 * abs(a) for a, b in zip(*losses))
    assert all(relnorm(b, a) <= 1e-5 for a, b in zip(*params))


def test_admm_dual_update_runs_for_ragged_last_batch_and_partial_forward():
    """with the AdmmBank (d trans_loss / d(alterD, gamma) not computed) the Z/U update must
    still happen when the batch differs from args.train_batch_size (the last partial batch of an epoch) and when the
    bank is not 'ready' -- the reference always updates (optimizer.py:97-124).  Reference here: the same step with
    fast_admm=False (per-module ADMM_OPT.step on autograd-visited parameters)."""
    Btrain = 16
    torch.manual_seed(6)
    xs = {b: torch.randn(b, 3, 32, 32, device=DEV) for b in (Btrain, 11)}
    ts = {b: torch.randint(0, 10, (b,), device=DEV) for b in (Btrain, 11)}
    res = {}
    for fast in (False, True):
        aq.reset_args()
        aq.set_args(variant="B", train_batch_size=Btrain, bitW=8, abitW=8, act_range=2, method="ours", gram_mode="fp32")
        m = resnet.resnet20_quant(8, 8, "second")
        m.load_state_dict(MO.deterministic_fill(m.state_dict(), seed=7))
        m.to(DEV).train()
        st = QATStep(m, fast_admm=fast, bank_weights=False)
        assert (st.admm_bank is not None) == fast
        assert aq.args.admm_param_grads is True                   # never mutated process-wide
        z0 = m.admm0.alterD.detach().clone()
        st.step(xs[Btrain], ts[Btrain])                           # ONE update at each batch size (the loop is chaotic)
        z1 = m.admm0.alterD.detach().clone()
        u1 = m.layers[4].admm1.gamma.detach().clone()
        st.step(xs[11], ts[11])                                   # ragged last batch: B = 11 < dim = 16
        res[fast] = (z0, z1, u1, m.admm0.alterD.detach().clone(), m.layers[4].admm1.gamma.detach().clone(),
                     m.layers[8].admm0.alterD.detach().clone())
        if fast:
            assert m.admm0.alterD.grad is None                    # closed-form update, no autograd gradient needed
            assert st.admm_bank.ready() == 11
            # bank not ready (a module's D handle is gone): the per-module fallback must still update Z/U
            zb = m.admm0.alterD.detach().clone()
            st.admm_bank.begin_iteration()
            out, tl = m(xs[11])
            m.layers[8].admm1.D = None
            assert st.admm_bank.ready() == 0
            from alignq_b200.utils.train import collect_admm_args
            st.opt_admm.step(*collect_admm_args(m, st.admm_params))
            assert not torch.equal(m.admm0.alterD.detach(), zb)
    for k in (1, 2, 3, 4, 5):
        assert not torch.equal(res[True][k], res[True][0]) or k in (2, 4, 5)
    assert not torch.equal(res[True][3], res[True][1]), "ragged batch: Z was not updated"
    assert relnorm(res[True][1], res[False][1]) <= 1e-5 and relnorm(res[True][2], res[False][2]) <= 1e-5
    # the second step starts from weights that already differ by round-off and the quantised net amplifies that
    # (see test_training_iterations_vs_oracle_trainer_on_gpu): same band as the other model-level checks
    assert relnorm(res[True][3], res[False][3]) <= 2e-2 and relnorm(res[True][4], res[False][4]) <= 2e-2
    assert relnorm(res[True][5], res[False][5]) <= 2e-2


@pytest.mark.parametrize("variant", ["A", "B"])
def test_driver_trains_checkpoints_and_resumes(variant, tmp_path):
    """epoch loop + MultiStepLR stepped by epoch number + the reference's checkpoint keys
    (main.py:125-151; cdf_alignment_admm/.../main.py:138-150) + --resume (main.py:101-113)."""
    from alignq_b200.utils.driver import Trainer
    B = 16
    aq.reset_args()
    aq.set_args(variant=variant, train_batch_size=B, bitW=8, abitW=8, act_range=2, method="ours", gram_mode="tf32x3")
    g = torch.Generator().manual_seed(0)
    train = [(torch.randn(B, 3, 32, 32, generator=g), torch.randint(0, 10, (B,), generator=g)) for _ in range(3)]
    train.append((torch.randn(11, 3, 32, 32, generator=g), torch.randint(0, 10, (11,), generator=g)))     # ragged last batch
    # eval batches must not exceed the ADMM dim = train_batch_size (reference quirk)
    test = [(torch.randn(12, 3, 32, 32, generator=g), torch.randint(0, 10, (12,), generator=g)) for _ in range(2)]

    def build():
        m = resnet.resnet20_quant(8, 8, "second")
        m.load_state_dict(MO.deterministic_fill(m.state_dict(), seed=21))
        return m.to(DEV)
    cfg = dict(job_dir=str(tmp_path / "run"), num_epochs=2, lr=0.04, momentum=0.9, weight_decay=1e-4, lr_decay_steps=[1],
               lr_gamma=0.1, print_freq=0)
    tr = Trainer(build(), cfg)
    hist = tr.fit(train, test)
    assert [h["lr"] for h in hist] == [0.04, pytest.approx(0.004)]                   # s.step(epoch): decayed AT epoch 1
    ck1 = torch.load(tmp_path / "run" / "checkpoint" / "model_1.pt", map_location=DEV, weights_only=False)
    want = {"state_dict_t", "best_prec1", "best_prec5", "optimizer_t", "scheduler_t", "epoch"} | ({"optimizer_admm"} if variant == "B" else set())
    assert set(ck1) == want and ck1["epoch"] == 1
    assert (tmp_path / "run" / "checkpoint" / "model_2.pt").exists() and (tmp_path / "run" / "checkpoint" / "model_best.pt").exists()
    keys = json.load(open(os.path.join(GOLDEN, f"model_keys_resnet20_{variant}.json")))["state_dict"]
    assert list(ck1["state_dict_t"].keys()) == list(keys)                                  # loadable by the reference's model
    st0 = ck1["optimizer_t"]["state"]
    assert all(set(v.keys()) == {"momentum_buffer"} for v in st0.values()) and len(st0) > 0
    if variant == "B":                                                                # Z/U moved (ragged batch included)
        fresh = build()
        assert not torch.equal(ck1["state_dict_t"]["admm0.alterD"], fresh.state_dict()["admm0.alterD"])
    # resume from epoch 1: identical restored state, then one more epoch at the decayed learning rate
    tr2 = Trainer(build(), dict(cfg, resume=str(tmp_path / "run" / "checkpoint" / "model_1.pt"), job_dir=str(tmp_path / "run2")))
    assert tr2.start_epoch == 1 and tr2.best_prec1 == ck1["best_prec1"]
    sd2 = tr2.model.state_dict()
    assert all(torch.equal(sd2[k], ck1["state_dict_t"][k]) for k in sd2)
    p0 = tr2.step.params[0]
    assert torch.equal(tr2.optimizer_t.state[p0]["momentum_buffer"], list(st0.values())[0]["momentum_buffer"].view_as(p0))
    hist2 = tr2.fit(train, test)
    assert len(hist2) == 1 and hist2[0]["epoch"] == 1 and hist2[0]["lr"] == pytest.approx(0.004)
    # the resumed epoch tracks the uninterrupted one (same data, same restored state; cuDNN wgrad order may differ)
    assert abs(hist2[0]["train_loss"] - hist[1]["train_loss"]) <= 2e-2 * abs(hist[1]["train_loss"])
