"""CPU: the model-level banks only re-point storage -- parameters keep their names, shapes, strides and values
and become views of one flat buffer (no kernels are launched at construction)."""
import torch

import alignq_b200 as aq
from alignq_b200.model import resnet
from alignq_b200.utils.admm_bank import AdmmBank
from alignq_b200.utils.weight_bank import WeightBank


def test_weight_bank_repoints_parameters_without_changing_them():
    aq.set_args(variant="A", bitW=8, abitW=8)
    torch.manual_seed(0)
    m = resnet.resnet20_quant(8, 8, "second")
    for p in m.parameters():                                   # channels_last weights must stay channels_last
        if p.dim() == 4:
            p.data = p.data.contiguous(memory_format=torch.channels_last)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    strides = {k: v.stride() for k, v in m.named_parameters()}
    bank = WeightBank(m)
    after = m.state_dict()
    assert list(after.keys()) == list(before.keys())
    assert all(torch.equal(after[k], before[k]) for k in before)
    assert all(v.stride() == strides[k] for k, v in m.named_parameters())
    assert len(bank.params) == 21 and bank.flat.numel() == sum(p.numel() for p in bank.params)
    lo, hi = bank.flat.data_ptr(), bank.flat.data_ptr() + 4 * bank.flat.numel()
    assert all(lo <= p.data_ptr() < hi for p in bank.params)
    with torch.no_grad():
        bank.flat.zero_()
    assert all(float(p.abs().max()) == 0.0 for p in bank.params)       # really views
    assert all(q._bank == (bank, i) for i, q in enumerate(bank.fns))
    assert bank.wq[3].shape == bank.params[3].shape and bank.wq[3].stride() == bank.params[3].stride()
    bank.release()
    assert all(q._bank is None for q in bank.fns)


def test_admm_bank_stacks_dual_variables():
    aq.set_args(variant="B", bitW=8, abitW=8, train_batch_size=6)
    torch.manual_seed(0)
    m = resnet.resnet20_quant(8, 8, "second")
    before = {k: v.clone() for k, v in m.state_dict().items()}
    bank = AdmmBank(m, batch=6)
    after = m.state_dict()
    assert list(after.keys()) == list(before.keys()) and all(torch.equal(after[k], before[k]) for k in before)
    assert bank.Z.shape == (len(bank.mods), 6, 6) and len(bank.mods) == 21
    assert m.admm0.alterD.data_ptr() == bank.Z[0].data_ptr()
    assert m.layers[0].act_q0.opt.alterD.data_ptr() == m.layers[0].admm0.alterD.data_ptr()   # shared module, shared view
    assert not bank.ready()                                     # no forward has written a D yet
    # the closed-form update is switched on PER MODULE: the process-global flag is untouched
    assert aq.args.admm_param_grads is True and all(mod.param_grads is False for mod in bank.mods)
    assert m.admm0.alterD._alignq_closed_form and m.admm0.gamma._alignq_closed_form
    # ragged last batch: D slots exist for any B <= dim, one [L, B, B] buffer per batch size
    assert m.admm0.d_slot(6).data_ptr() == bank.D[0].data_ptr()
    s4 = m.admm0.d_slot(4)
    assert s4.shape == (4, 4) and m.layers[0].admm0.d_slot(4).data_ptr() != s4.data_ptr()
    try:
        m.admm0.d_slot(7)
        raise AssertionError("batch above dim must raise")
    except aq.AlignQError:
        pass
    bank.release()
    assert all(mod.param_grads is True and mod._bank is None for mod in bank.mods)
    assert not hasattr(m.admm0.alterD, "_alignq_closed_form")
