#!/usr/bin/env python
"""Build a full model of a workload structure THROUGH THE REFERENCE's own construction path
(``ReferenceStructure`` -> ``InterpolationModel.add_dof`` / ``ARTModel.add_art``), with this package's
space-group search in place of spglib and the vectorised basis-vector scans of ``construction.py``
(SURVEY.md §8f row N4), and report the wall time.  Needs an importable reference tree (this container).

    python tools/build_model_reference.py LLZO_2x2x2 [--art] [--no-accelerate]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_bootstrap import import_reference  # noqa: E402
from ramannoodle_b200 import construction, symmetry, synthetic  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("structure", nargs="?", default="LLZO_2x2x2")
    ap.add_argument("--art", action="store_true", help="ARTModel.add_art instead of InterpolationModel.add_dof")
    ap.add_argument("--no-accelerate", action="store_true", help="keep the reference's Python scans")
    args = ap.parse_args()
    import_reference()
    import ramannoodle.structure._reference as ref_module
    from ramannoodle.exceptions import InvalidDOFException
    from ramannoodle.pmodel._art import ARTModel
    from ramannoodle.pmodel._interpolation import InterpolationModel

    ref_module.spglib.get_symmetry = symmetry.get_symmetry
    if not args.no_accelerate:
        construction.accelerate_construction()
    data = synthetic.load_structure(args.structure)
    times = {}
    t0 = time.perf_counter()
    structure = ref_module.ReferenceStructure([int(z) for z in data["atomic_numbers"]], data["lattice"], data["positions"])
    times["reference_structure_s"] = round(time.perf_counter() - t0, 2)
    rng = np.random.default_rng(1)
    ref_polarizability = np.diag([6.0, 6.0, 6.5])
    model = (ARTModel if args.art else InterpolationModel)(structure, ref_polarizability)
    num_atoms = len(data["positions"])
    calls = rejected = 0
    t0 = time.perf_counter()
    for atom in sorted(structure.get_equivalent_atom_dict()):
        for axis in np.eye(3):
            if len(model.cart_basis_vectors) == 3 * num_atoms:
                break
            # the part of this axis that the DOFs already on this atom do not cover (every DOF moves one atom)
            direction = axis.copy()
            for vector in model.cart_basis_vectors:
                row = vector[atom]
                norm = np.linalg.norm(row)
                if norm > 0.5:  # unit vectors: the atom this DOF moves (other rows hold symmetry-tolerance noise)
                    direction -= (direction @ row) / norm ** 2 * row
            if np.linalg.norm(direction) < 1e-6:
                continue
            direction /= np.linalg.norm(direction)

            def tensor():
                t = rng.normal(scale=0.05, size=(3, 3))
                return ref_polarizability + 0.5 * (t + t.T)
            displacement = np.zeros((num_atoms, 3))
            displacement[atom] = direction @ np.linalg.inv(data["lattice"])
            displacement /= np.linalg.norm(structure.get_cart_displacement(displacement))
            # amplitudes related by symmetry must be given once (the reference refuses the redundant ones)
            for amplitudes in ((np.array([-0.1, 0.1]), np.array([0.1])) if args.art else
                               (np.array([-0.1, -0.05, 0.05, 0.1]), np.array([0.05, 0.1]))):
                try:
                    values = np.array([tensor() for _ in amplitudes])
                    if args.art:
                        model.add_art(atom, direction, amplitudes, values)
                    else:
                        model.add_dof(displacement, amplitudes, values, 3 if len(amplitudes) > 2 else 2)
                    calls += 1
                    break
                except InvalidDOFException as exc:
                    rejected += 1
                    if os.environ.get("RN_VERBOSE"):
                        print(f"atom {atom} direction {np.round(direction, 4)} x{len(amplitudes)}: {exc}", file=sys.stderr)
    times["add_dof_s"] = round(time.perf_counter() - t0, 2)
    print(json.dumps({"structure": args.structure, "atoms": num_atoms, "operations": len(structure._rotations),  # pylint: disable=protected-access
                      "nonequivalent_atoms": structure.num_nonequivalent_atoms, "dofs": len(model.cart_basis_vectors),
                      "calls": calls, "rejected": rejected, "accelerated_scans": not args.no_accelerate,
                      "kind": "art" if args.art else "cubic", **times}))


if __name__ == "__main__":
    main()
