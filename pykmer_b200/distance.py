"""distance: .kma matrix -> Jaccard distance matrix -> neighbour-joining tree.

Host-side mirror of the reference's downstream consumer calculate_distance.py (SURVEY.md 8f,
rank 3): it defines which cells of the merger's (N, N, 3) matrix matter.  This is O(N^2) float
work on at most 255 x 255 cells, so it stays on the host in NumPy -- no kernel, not a hot path.

  get_matrix        calculate_distance.py:29-40    np.load(<x.kma>)["matrix"]
  calc_distance     calculate_distance.py:42-109   dist = 1 - shared / (totalA + totalB - shared)
                                                   (:82-84), diagonal := 0 (:96-97), saved as
                                                   <x.kma>.dist.jaccard.npz (key "distance")
  cluster_distance  calculate_distance.py:111-235  ids from the .kma.json (input_file_name per sample,
                                                   :137-146) mapped through an optional names file
                                                   (:21-27,148-150); writes
                                                     .mat.redundant.np / .mat.redundant.lsmat
                                                     .mat.condensed.np / .mat.condensed.txt
                                                     .newick / .tree
  load, main        calculate_distance.py:237-249

Parity: calc_distance and the four matrix dumps are pinned against the reference's own
calc_distance run in the build container (oracle/make_golden_distance.py ->
tests/golden/distance/).  The tree is NOT pinned: the reference delegates it to scikit-bio's
`nj` and ete3's renderer, neither of which is installed here, so `neighbor_joining` restates
the published algorithm (Saitou & Nei 1987 in the Studier & Keppler Q-matrix form, negative
branch lengths clamped to 0 as scikit-bio does by default) and the .png rendering is left out.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def read_names_file(names_file: Path) -> Dict[str, str]:
    """file name -> display name, two tab-separated columns (calculate_distance.py:21-27)."""
    names_file = Path(names_file)
    assert names_file.exists()
    names = {}
    with names_file.open("rt") as fhd:
        for row in fhd:
            cols = row.split("\t")
            if len(cols) == 2:
                names[cols[0].strip()] = cols[1].strip()
    return names


def get_matrix(matrix_file: Path) -> np.ndarray:
    matrix_file = Path(matrix_file)
    assert matrix_file.exists() and matrix_file.is_file()
    npz = np.load(matrix_file)
    assert "matrix" in npz
    return npz["matrix"]


def jaccard_distance(matrix: np.ndarray, fill_diagonal: bool = True) -> np.ndarray:
    """1 - shared / (totalA + totalB - shared) per cell (calculate_distance.py:82-84,96-97)."""
    shared = matrix[:, :, 2].astype(np.float64)
    total = matrix[:, :, 0:2].sum(axis=2).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        dist = 1.0 - (shared / (total - shared))
    if fill_diagonal:
        np.fill_diagonal(dist, 0.0)
    return dist


def calc_distance(matrix_file: Path, matrix: np.ndarray, fill_diagonal: bool = True) -> Tuple[Path, np.ndarray]:
    dist = jaccard_distance(matrix, fill_diagonal)
    basefile = Path(f"{matrix_file}.dist.jaccard")
    with Path(f"{basefile}.npz").open("wb") as fhd:
        np.savez(fhd, distance=dist)
    return basefile, dist


def condensed_form(distance: np.ndarray) -> np.ndarray:
    """Upper triangle, row by row (what scikit-bio's DistanceMatrix.condensed_form returns)."""
    iu = np.triu_indices(distance.shape[0], k=1)
    return np.ascontiguousarray(distance[iu])


def write_lsmat(fhd, distance: np.ndarray, ids: Sequence[str], delimiter: str = "\t") -> None:
    """Labelled square matrix: a header row of ids, then one row per id."""
    fhd.write(delimiter + delimiter.join(ids) + "\n")
    for name, row in zip(ids, distance):
        fhd.write(delimiter.join([name] + [repr(float(v)) for v in row]) + "\n")


def neighbor_joining(distance: np.ndarray, ids: Sequence[str], clamp_negative: bool = True) -> str:
    """Neighbour-joining tree of a symmetric distance matrix as a Newick string (unrooted,
    trifurcating at the last join)."""
    n = distance.shape[0]
    assert distance.shape == (n, n) and len(ids) == n
    if n < 3:
        raise ValueError("neighbour joining needs at least 3 taxa")
    d = np.array(distance, dtype=np.float64)
    labels: List[str] = [str(i) for i in ids]

    def limb(v: float) -> float:
        return max(v, 0.0) if clamp_negative else v

    while d.shape[0] > 3:
        m = d.shape[0]
        r = d.sum(axis=1)
        q = (m - 2) * d - r[:, None] - r[None, :]
        np.fill_diagonal(q, np.inf)
        i, j = divmod(int(np.argmin(q)), m)
        if i > j:
            i, j = j, i
        li = limb(0.5 * d[i, j] + (r[i] - r[j]) / (2.0 * (m - 2)))
        lj = limb(d[i, j] - li)
        node = f"({labels[i]}:{li:.6f}, {labels[j]}:{lj:.6f})"
        du = 0.5 * (d[i, :] + d[j, :] - d[i, j])
        if clamp_negative:
            du = np.maximum(du, 0.0)
        keep = [k for k in range(m) if k not in (i, j)]
        nd = np.zeros((m - 1, m - 1), dtype=np.float64)
        nd[0, 1:] = du[keep]
        nd[1:, 0] = du[keep]
        nd[1:, 1:] = d[np.ix_(keep, keep)]
        d = nd
        labels = [node] + [labels[k] for k in keep]
    l0 = limb(0.5 * (d[0, 1] + d[0, 2] - d[1, 2]))
    l1 = limb(0.5 * (d[0, 1] + d[1, 2] - d[0, 2]))
    l2 = limb(0.5 * (d[0, 2] + d[1, 2] - d[0, 1]))
    return f"({labels[0]}:{l0:.6f}, {labels[1]}:{l1:.6f}, {labels[2]}:{l2:.6f});"


def _parse_newick(text: str):
    """Newick -> nested (label, length, children) tuples (enough for the trees written here)."""
    pos = 0

    def node():
        nonlocal pos
        children = []
        if text[pos] == "(":
            pos += 1
            while True:
                while text[pos] in " ,":
                    pos += 1
                children.append(node())
                while text[pos] == " ":
                    pos += 1
                if text[pos] == ")":
                    pos += 1
                    break
        start = pos
        while pos < len(text) and text[pos] not in ":,); ":
            pos += 1
        label = text[start:pos]
        length = None
        if pos < len(text) and text[pos] == ":":
            pos += 1
            start = pos
            while pos < len(text) and text[pos] not in ",); ":
                pos += 1
            length = float(text[start:pos])
        return label, length, children

    return node()


def ascii_tree(newick: str) -> str:
    """A plain text drawing of the tree (the reference writes ete3's str(Tree), :211-214)."""
    root = _parse_newick(newick.strip())
    lines: List[str] = []

    def draw(n, prefix: str, tail: bool, top: bool):
        label, _, children = n
        if not top:
            lines.append(prefix + ("\\-" if tail else "|-") + (label if not children else "+"))
        else:
            lines.append("+")
        for k, c in enumerate(children):
            draw(c, prefix + ("" if top else ("  " if tail else "| ")), k == len(children) - 1, False)

    draw(root, "", True, True)
    return "\n".join(lines) + "\n"


def cluster_distance(matrix_file: Path, basefile: Path, distance: np.ndarray, names_file: Optional[Path] = None,
                     load_header: bool = True, save_matrix_redundant_tsv: bool = True,
                     save_matrix_redundant_np: bool = True, save_matrix_condensed_tsv: bool = True,
                     save_matrix_condensed_np: bool = True, save_tree_newick: bool = True,
                     save_tree_ascii: bool = True) -> np.ndarray:
    if load_header:
        with Path(f"{matrix_file}.json").open("rt") as fhd:
            header = json.load(fhd)
        ids = [d["header"]["input_file_name"] for d in header["data"]]
        assert len(ids) == distance.shape[0]
    else:
        ids = [str(d + 1) for d in range(distance.shape[0])]
    if names_file:
        names = read_names_file(names_file)
        ids = [names.get(i, i) for i in ids]
    # what scikit-bio's DistanceMatrix would insist on (calculate_distance.py:153)
    assert len(set(ids)) == len(ids), "sample ids must be unique"
    assert np.allclose(distance, distance.T, equal_nan=True), "distance matrix must be symmetric"

    dmr = np.ascontiguousarray(distance, dtype=np.float64)
    if save_matrix_redundant_np:
        with Path(f"{basefile}.mat.redundant.np").open("wb") as fhd:
            np.save(fhd, dmr, allow_pickle=False)
    if save_matrix_redundant_tsv:
        with Path(f"{basefile}.mat.redundant.lsmat").open("wt") as fhd:
            write_lsmat(fhd, dmr, ids)
    if save_matrix_condensed_np or save_matrix_condensed_tsv:
        dmc = condensed_form(dmr)
        if save_matrix_condensed_np:
            with Path(f"{basefile}.mat.condensed.np").open("wb") as fhd:
                np.save(fhd, dmc, allow_pickle=False)
        if save_matrix_condensed_tsv:
            with Path(f"{basefile}.mat.condensed.txt").open("wt") as fhd:
                np.savetxt(fhd, dmc)
    if (save_tree_newick or save_tree_ascii) and distance.shape[0] >= 3:
        newick = neighbor_joining(dmr, ids)
        if save_tree_newick:
            Path(f"{basefile}.newick").write_text(newick)
        if save_tree_ascii:
            Path(f"{basefile}.tree").write_text(ascii_tree(newick))
    return dmr


def load(matrix_file: Path, names_file: Optional[Path] = None) -> np.ndarray:
    matrix_file = Path(matrix_file)
    if names_file is None:
        candidate = Path(f"{matrix_file}.names.tsv")
        if candidate.exists():
            names_file = candidate
    matrix = get_matrix(matrix_file)
    basefile, distance = calc_distance(matrix_file, matrix, fill_diagonal=True)
    return cluster_distance(matrix_file, basefile, distance, names_file=names_file)


def main(argv: Optional[Sequence[str]] = None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) != 1:
        print("usage: calculate_distance.py <project.MIN-MAX.kma>", file=sys.stderr)
        sys.exit(1)
    load(Path(argv[0]))


if __name__ == "__main__":
    main()
