"""TEST / BASELINE INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference (aiqm/torchani).

``oracle/build_ref.sh`` stages the reference's Python package together with its own compiled extensions
(cuAEV, cell_list, MNP; sm_100) into the git-ignored ``oracle/_ref/torchani/``; that directory travels to
the GPU box (``/root/reference`` does not).  ``load()`` imports it -- never anything of this repo's product --
and ``build_model()`` assembles an ANI-1x/2x shaped ``torchani.arch.ANI`` (arch.py:300-381) holding the
seeded synthetic weights every test and benchmark of this repo uses, with the reference's own choice of

* AEV strategy: ``"pyaev"`` (pure PyTorch, the CPU path) or ``"cuaev"`` (its CUDA extension,
  aev/_computer.py:383-407),
* neighbour list: ``"cell_list"`` / ``"all_pairs"`` (neighbors.py),
* networks: the python ``Ensemble`` loop, or ``.to_infer_model()`` = ``BmmEnsemble`` (nn/_infer.py:61-216) /
  MNP (``use_mnp=True``).

Users: tests/ (drop-in and parity tests), bench.py's ``--impl reference`` / ``cpu_baseline`` legs (the thing
timed beside the product, never the product) and tools/reference_gpu_path.py.
"""
from __future__ import annotations

import os
import sys
import types
import typing as tp
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
SOURCE = "/root/reference"

_mod = None


def available() -> bool:
    return os.path.isdir(os.path.join(STAGED, "torchani")) or os.path.isdir(os.path.join(SOURCE, "torchani"))


def load():
    """Import the reference package: the staged copy (with its compiled extensions) if present, else the
    source tree of the build container (pure-PyTorch paths only).  h5py is absent from the image and only
    needed by the reference's dataset code: a three-attribute stub lets ``import torchani`` succeed
    (SURVEY.md 8c)."""
    global _mod
    if _mod is not None:
        return _mod
    root = STAGED if os.path.isdir(os.path.join(STAGED, "torchani")) else SOURCE
    if not os.path.isdir(os.path.join(root, "torchani")):
        raise ImportError("the reference is not staged: run oracle/build_ref.sh where /root/reference exists")
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            stub = types.ModuleType("h5py")
            stub.Group = stub.File = stub.Dataset = type("_Stub", (), {})
            sys.modules["h5py"] = stub
    os.environ.setdefault("TORCHANI_NO_WARN_EXTENSIONS", "1")
    if root not in sys.path:
        sys.path.insert(0, root)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import torchani  # noqa: F401
    _mod = torchani
    return torchani


def extensions() -> tp.Dict[str, bool]:
    load()
    from torchani.csrc import CLIST_IS_INSTALLED, CUAEV_IS_INSTALLED, MNP_IS_INSTALLED
    return {"cuaev": bool(CUAEV_IS_INSTALLED), "cell_list": bool(CLIST_IS_INSTALLED), "mnp": bool(MNP_IS_INSTALLED)}


GSAES_WB97X_631GD = {"H": -0.4993212, "C": -37.8338334, "N": -54.5732825, "O": -75.0424519,
                     "S": -398.0814169, "F": -99.6949007, "Cl": -460.1167008}   # constants.py:88-96


def build_model(weights, kind: str = "2x", device="cpu", dtype: torch.dtype = torch.float32,
                strategy: str = "pyaev", neighborlist: str = "cell_list", infer: bool = False,
                use_mnp: bool = False, periodic_table_index: bool = True, repulsion: bool = False):
    """A real ``torchani.arch.ANI`` with ``weights[member][symbol] = [(W [out,in], b [out]) x 4]``."""
    ta = load()
    from torchani.aev import AEVComputer
    from torchani.arch import ANI
    from torchani.nn import ANINetworks, Ensemble
    from torchani.sae import SelfEnergy

    symbols = ("H", "C", "N", "O", "S", "F", "Cl") if kind == "2x" else ("H", "C", "N", "O")
    mk = AEVComputer.like_2x if kind == "2x" else AEVComputer.like_1x
    aevc = mk(neighborlist=neighborlist, strategy=strategy)
    members = []
    for w_m in weights:
        net = ANINetworks.like_2x() if kind == "2x" else ANINetworks.like_1x()
        with torch.no_grad():
            for s in symbols:
                an = net.atomics[s]
                for lin, (w, b) in zip(list(an.layers) + [an.final_layer], w_m[s]):
                    assert lin.weight.shape == w.shape, (s, lin.weight.shape, w.shape)
                    lin.weight.copy_(w)
                    lin.bias.copy_(b)
        members.append(net)
    nets = Ensemble(members) if len(members) > 1 else members[0]
    extra = None
    if repulsion:   # the pair potential of ANI-2xr (models.py:255-290): shares the model's neighbour list
        from torchani.potentials import RepulsionXTB
        extra = {"repulsion_xtb": RepulsionXTB(symbols, cutoff=aevc.radial.cutoff)}
    model = ANI(symbols, aevc, nets, SelfEnergy(symbols, [GSAES_WB97X_631GD[s] for s in symbols]),
                potentials=extra, periodic_table_index=periodic_table_index)
    model = model.to(device=device, dtype=dtype)
    model.requires_grad_(False)
    if infer:
        model = model.to_infer_model(use_mnp=use_mnp)
    del ta
    return model


def energies_and_forces(model, species, coords, cell=None, pbc=None):
    """grad.py:263-290 (``torchani.grad.energies_and_forces``): the reference's own energy+force call."""
    load()
    from torchani.grad import energies_and_forces as eaf
    out = eaf(model, species, coords, cell, pbc)
    return out.energies, out.forces
