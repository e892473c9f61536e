"""TEST INFRASTRUCTURE ONLY -- a restatement of the reference's inference-optimised network path on the GPU
(``torchani.nn.BmmEnsemble``, nn/_infer.py:61-216): per element one ``torch.baddbmm`` per layer over the
stacked weights of all ensemble members, CELU(0.1) in between, mean over the members, sum over the atoms.
Plain PyTorch / cuBLAS: this is what the reference's GPU path runs for the networks when MNP is not used.
Only tests/ import it (as the thing that is timed and compared beside the B200 kernels)."""
from __future__ import annotations

import typing as tp

import torch
from torch import Tensor


class BmmNetworks:
    """weights[member][symbol] = [(W [out, in], b [out]) x 4] as in oracle.ani_oracle.make_weights."""

    def __init__(self, weights: tp.Sequence[tp.Mapping[str, tp.Sequence[tp.Tuple[Tensor, Tensor]]]],
                 symbols: tp.Sequence[str], device: torch.device, dtype: torch.dtype = torch.float32):
        self.symbols = tuple(symbols)
        self.layers: tp.List[tp.List[tp.Tuple[Tensor, Tensor]]] = []
        for s in self.symbols:
            per_layer = []
            for k in range(len(weights[0][s])):
                # BmmLinear (nn/_infer.py:171-203): weight (e, in, out), bias (e, 1, out)
                w = torch.stack([wm[s][k][0].t() for wm in weights]).to(device=device, dtype=dtype).contiguous()
                b = torch.stack([wm[s][k][1].view(1, -1) for wm in weights]).to(device=device, dtype=dtype).contiguous()
                per_layer.append((w, b))
            self.layers.append(per_layer)

    def make_idx_list(self, elem_idxs: Tensor) -> tp.List[Tensor]:
        flat = elem_idxs.flatten()
        return [(flat == i).nonzero().flatten() for i in range(len(self.symbols))]   # nn/_infer.py:42-58

    def energy(self, elem_idxs: Tensor, aevs: Tensor, idx_list: tp.Optional[tp.List[Tensor]] = None) -> Tensor:
        """BmmEnsemble.forward (nn/_infer.py:103-129) + BmmAtomicNetwork.forward (:163-167): (1,) energy."""
        assert aevs.shape[0] == 1, "single-conformer inputs only, as the reference"
        x = aevs.flatten(0, 1)
        if idx_list is None:
            idx_list = self.make_idx_list(elem_idxs)
        energies = x.new_zeros(x.shape[0])
        for i, layers in enumerate(self.layers):
            if idx_list[i].shape[0] == 0:
                continue
            h = x.index_select(0, idx_list[i]).expand(layers[0][0].shape[0], -1, -1)
            for w, b in layers[:-1]:
                h = torch.nn.functional.celu(torch.baddbmm(b, h, w), alpha=0.1)
            w, b = layers[-1]
            energies = energies.index_put((idx_list[i],), torch.baddbmm(b, h, w).mean(0).flatten())
        return energies.view_as(elem_idxs).sum(dim=-1)
