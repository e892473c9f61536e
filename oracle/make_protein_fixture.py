"""Fixtures for BASELINE config 5 (protein in water) from the reference's dataset/pdb/1C17.pdb.

TEST INFRASTRUCTURE ONLY; runs in the build container (the GPU box has no /root/reference).  Writes

* tests/golden/1c17_protein.npz -- element indices (ANI-2x order H C N O S F Cl) and coordinates (float32, A)
  of the 16 649 protein atoms: the solute of `torchani_b200.synthetic.protein_in_water`, the 50k-atom
  workload of `bench.py --config protein50k`;
* tests/golden/1c17_chunk_hcnos_ani2x.npz -- a golden parity case with ALL FIVE elements of the protein
  (H C N O S: 20 of the 32 AEV column blocks live, against 5 of 32 for water): the atoms within 13 A of the
  first sulfur, no PBC, evaluated with the REAL reference (pure-PyTorch pyaev path, float64, cell_list) and
  pinned against the oracle exactly like the cases of oracle/make_golden.py.

Usage:  python oracle/make_protein_fixture.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
PDB = "/root/reference/dataset/pdb/1C17.pdb"


def read_pdb(path):
    sym, xyz = [], []
    for line in open(path):
        if line.startswith(("ATOM", "HETATM")):
            sym.append(line[76:78].strip().capitalize())
            xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return sym, np.array(xyz, dtype=np.float64)


def main():
    import oracle.ani_oracle as orc
    from oracle.make_golden import build_reference, import_reference, run_reference

    sym, xyz = read_pdb(PDB)
    idx_of = {s: i for i, s in enumerate(orc.SYMBOLS_2X)}
    species = np.array([idx_of[s] for s in sym], dtype=np.uint8)
    out = os.path.join(ROOT, "tests", "golden")
    np.savez_compressed(os.path.join(out, "1c17_protein.npz"), species=species, coords=xyz.astype(np.float32),
                        source="aiqm/torchani dataset/pdb/1C17.pdb (ATOM records, element column)")
    print("1c17_protein:", len(sym), "atoms", {s: int((species == i).sum()) for s, i in idx_of.items()})

    # ---- golden chunk with all five elements
    s0 = int(np.nonzero(species == idx_of["S"])[0][0])
    keep = np.nonzero(np.linalg.norm(xyz - xyz[s0], axis=1) <= 13.0)[0]
    c_idx = torch.tensor(species[keep].astype(np.int64)).unsqueeze(0)
    c_xyz = torch.tensor(xyz[keep] - xyz[keep].mean(0)).unsqueeze(0)
    print("chunk:", len(keep), "atoms", {s: int((c_idx == i).sum()) for s, i in idx_of.items()})
    torchani = import_reference()
    w2 = orc.make_weights(orc.SYMBOLS_2X, orc.DIMS_2X, 1008, 8, 1234, torch.float64)
    aevc, ens = build_reference(torchani, "2x", w2, "cell_list", orc.SYMBOLS_2X, orc.DIMS_2X)
    ref = run_reference(aevc, ens, c_idx, c_xyz, None, None)
    sae = {s_: orc.GSAES_WB97X_631GD[s_] for s_ in orc.SYMBOLS_2X}
    model = orc.Model(orc.aev_spec_2x(), orc.SYMBOLS_2X, w2, sae, "cell_list")
    mine = orc.compute(model, c_idx, c_xyz, None, None)
    errs = {k: float((mine[k] - ref[k]).abs().max()) for k in ("aev", "member_atomic", "forces")}
    assert int(mine["num_pairs"]) == ref["num_pairs"] and max(errs.values()) < 1e-11, errs
    print("oracle-vs-reference max-abs", errs, "pairs", ref["num_pairs"])
    proj = np.random.default_rng(99).standard_normal((1008, 4))
    np.savez_compressed(
        os.path.join(out, "1c17_chunk_hcnos_ani2x.npz"), kind="2x", neighborlist="cell_list", weight_seed=1234,
        species=c_idx.numpy(), coords=c_xyz.numpy(), cell=np.zeros((0, 3)), pbc=np.zeros(0, dtype=bool),
        num_pairs=ref["num_pairs"], member_atomic=ref["member_atomic"].numpy(), energy_nn=ref["energy_nn"].numpy(),
        forces=ref["forces"].numpy(), aev_proj_matrix=proj, aev_proj=ref["aev"].numpy() @ proj,
        aev_rowsum=ref["aev"].numpy().sum(-1))


if __name__ == "__main__":
    main()
