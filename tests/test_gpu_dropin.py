"""Drop-in proof against the REAL reference package (oracle/_ref/torchani, staged by oracle/build_ref.sh).

Two ways a TorchANI user switches to the B200 path, both exercised on a real ``torchani.arch.ANI``:

* ``torchani_b200.models.from_torchani(ref_model)``: convert the whole model (arch.py:90-381 -> models.ANI),
* ``torchani_b200.models.accelerate_torchani_(ref_model)``: swap the three plug-in points IN PLACE --
  ``ref.neighborlist``, ``ref.potentials["nnp"].aev_computer``, ``ref.potentials["nnp"].neural_networks``
  (arch.py:116-127, 208-217, 264-275) -- and keep calling the reference's own ``forward`` /
  ``compute_from_neighbors`` / ``torchani.grad.energies_and_forces``.

The yardstick is the unmodified reference itself (pure-PyTorch ``pyaev`` path, float64, on the CPU) on the same
inputs and weights; bars as everywhere: forces 1e-4 Ha/A, energies 1e-5 relative.  Skipped when the staged
reference is absent (it is built where /root/reference exists and travels to the GPU box).
"""
import os

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "torchani")):
    pytest.skip("oracle/_ref/torchani is not staged (oracle/build_ref.sh needs /root/reference)", allow_module_level=True)


def _weights(members=8, dtype=torch.float32):
    from torchani_b200 import models, synthetic
    return synthetic.make_weights(models.SYMBOLS_2X, synthetic.DIMS_2X, 1008, members, seed=1234, dtype=dtype)


def _reference_f64(weights64, z, coords, cell, pbc, neighborlist, **kw):
    import oracle.ref_torchani as rt
    ref = rt.build_model(weights64, "2x", "cpu", dtype=torch.float64, strategy="pyaev", neighborlist=neighborlist, **kw)
    c = coords.double().clone()
    e, f = rt.energies_and_forces(ref, z, c, None if cell is None else cell.double(), pbc)
    return e.detach(), f.detach()


def _water(n_mol):
    from torchani_b200 import synthetic
    z, _, coords, cell, pbc = synthetic.water_box(n_mol, seed=3)
    return z, coords, cell, pbc


def test_from_torchani_converts_a_real_reference_model():
    import oracle.ref_torchani as rt
    from torchani_b200 import models
    z, coords, cell, pbc = _water(100)
    e64, f64 = _reference_f64(_weights(dtype=torch.float64), z, coords, cell, pbc, "cell_list")
    ref32 = rt.build_model(_weights(), "2x", DEV, strategy="pyaev", neighborlist="cell_list")
    ours = models.from_torchani(ref32)
    assert isinstance(ours, models.ANI) and len(ours) == 8 and ours.periodic_table_index
    assert next(ours.buffers()).device.type == "cuda"
    c = coords.to(DEV).requires_grad_(True)
    out = ours((z.to(DEV), c), cell.to(DEV), pbc.to(DEV))
    (g,) = torch.autograd.grad(out.energies.sum(), c)
    assert float((-g.cpu().double() - f64).abs().max()) < 1e-4
    e_f64 = ours.energies_f64((z.to(DEV), coords.to(DEV)), cell.to(DEV), pbc.to(DEV))
    assert abs(float(e_f64[0]) - float(e64[0])) < 1e-5 * 300       # 1e-5 Ha per atom on |E| ~ 7.6e3 Ha
    # active members travel with the conversion
    ref32.set_active_members([1, 4])
    sub = models.from_torchani(ref32)
    assert sub.neural_networks.active_members_idxs == [1, 4] and len(sub) == 2
    # ... and so does the inference-optimised container (BmmEnsemble, nn/_infer.py:61-216)
    ref32.set_active_members(list(range(8)))
    bmm = models.from_torchani(ref32.to_infer_model())
    c2 = coords.to(DEV).requires_grad_(True)
    (g2,) = torch.autograd.grad(bmm((z.to(DEV), c2), cell.to(DEV), pbc.to(DEV)).energies.sum(), c2)
    assert float((g2 - g).abs().max()) < 1e-6


def test_module_swap_inside_a_real_reference_model_periodic():
    """The reference's own ANI.forward (arch.py:302-349, the neighbour-list branch) and
    torchani.grad.energies_and_forces run on top of the three swapped modules."""
    import oracle.ref_torchani as rt
    from torchani_b200 import aev, models, neighbors, nn
    z, coords, cell, pbc = _water(100)
    e64, f64 = _reference_f64(_weights(dtype=torch.float64), z, coords, cell, pbc, "cell_list")
    ref = rt.build_model(_weights(), "2x", DEV, strategy="pyaev", neighborlist="cell_list")
    ta = rt.load()
    assert isinstance(ref, ta.arch.ANI)
    models.accelerate_torchani_(ref)
    assert isinstance(ref.neighborlist, neighbors.CellList)
    assert isinstance(ref.aev_computer, aev.AEVComputer) and isinstance(ref.neural_networks, nn.Ensemble)
    e, f = rt.energies_and_forces(ref, z.to(DEV), coords.to(DEV), cell.to(DEV), pbc.to(DEV))
    assert float((f.cpu().double() - f64).abs().max()) < 1e-4
    assert abs(float(e[0].detach()) - float(e64[0])) < 5e-3     # float32 total at |E| ~ 7.6e3 Ha (reference arithmetic)
    # atomic energies / ensemble values through the reference's own entry points
    at = ref.atomic_energies((z.to(DEV), coords.to(DEV)), cell.to(DEV), pbc.to(DEV)).energies
    assert at.shape == (1, 300) and abs(float(at.sum()) - float(e[0])) < 5e-2
    ev = ref((z.to(DEV), coords.to(DEV)), cell.to(DEV), pbc.to(DEV), ensemble_values=True).energies
    assert ev.shape == (8, 1) and abs(float(ev.mean()) - float(e[0])) < 5e-3
    # set_active_members reaches the B200 container through the reference's method (arch.py:132-134)
    ref.set_active_members([0, 5])
    assert ref.neural_networks.active_members_idxs == [0, 5] and len(ref) == 2


def test_module_swap_inside_a_real_reference_model_nonperiodic_batch():
    """Non-periodic padded batch: the reference's ANI.forward takes its 'fused' branch and calls
    aev_computer(elem_idxs, coords, cell, pbc) directly (arch.py:317-331)."""
    import oracle.ref_torchani as rt
    from torchani_b200 import models, synthetic
    species, coords = synthetic.conformer_batch(16, seed=5)
    conv = torch.tensor([1, 6, 7, 8])
    z = torch.where(species >= 0, conv[species.clamp(min=0)], torch.full_like(species, -1))
    e64, f64 = _reference_f64(_weights(dtype=torch.float64), z, coords, None, None, "all_pairs")
    ref = rt.build_model(_weights(), "2x", DEV, strategy="pyaev", neighborlist="all_pairs")
    models.accelerate_torchani_(ref)
    e, f = rt.energies_and_forces(ref, z.to(DEV), coords.to(DEV))
    assert e.shape == (16,)
    assert float((f.cpu().double() - f64).abs().max()) < 1e-4
    assert float((e.cpu().double() - e64).abs().max()) < 2e-3      # float32 totals at |E| ~ 1e3 Ha
    pad = (z == -1)
    assert float(f.cpu()[pad].abs().max()) == 0.0


def test_reference_pair_potential_runs_on_the_swapped_neighbor_list():
    """SURVEY 8(f4) tail: pair potentials sharing the model's neighbour list.  A real reference model WITH an extra
    pair potential (RepulsionXTB, as in ANI-2xr, models.py:255-290, potentials/xtb.py) keeps working after the module
    swap: the reference's own compute_from_neighbors loop (arch.py:354-381) hands the B200 neighbour list to every
    potential (discard_outside_cutoff per potential) and sums the energies; forces by autograd through both."""
    import oracle.ref_torchani as rt
    from torchani_b200 import models
    z, coords, cell, pbc = _water(100)
    e64, f64 = _reference_f64(_weights(dtype=torch.float64), z, coords, cell, pbc, "cell_list", repulsion=True)
    e_plain, _ = _reference_f64(_weights(dtype=torch.float64), z, coords, cell, pbc, "cell_list")
    assert abs(float(e64[0]) - float(e_plain[0])) > 1e-3          # the repulsion term is really there
    ref = rt.build_model(_weights(), "2x", DEV, strategy="pyaev", neighborlist="cell_list", repulsion=True)
    models.accelerate_torchani_(ref)
    assert set(ref.potentials.keys()) == {"repulsion_xtb", "nnp"}
    e, f = rt.energies_and_forces(ref, z.to(DEV), coords.to(DEV), cell.to(DEV), pbc.to(DEV))
    scale = max(1.0, float(f64.abs().max()))
    assert float((f.cpu().double() - f64).abs().max()) < 1e-4 * scale
    assert abs(float(e[0].detach()) - float(e64[0])) < 5e-3
