"""The similarity / dissimilarity zoo over 2x2 presence tables -- host-side mirror of the
reference's calculate_distances_cnidaria.py (statholder :40-473, attachMethodName :475-549,
init :552-569).

For a pair of samples X, Y the merger's counts give the table

              Y present   Y absent
  X present       a           b          a = shared k-mers, b = countX - a, c = countY - a
  X absent        c           d          (B, C: the same with totalX / totalY)

and every method maps it to one number.  The reference keeps ONE mutable `stats_data` object, fills
it per pair and lets ~70 bound methods read it; names (`S_*` similarities reported as 1 - s,
`D_*` distances), the derived fields (aPb = a + b, aTd = a * d, ...), the odd definition
d = a + b + c (calculate_distances_cnidaria.py:499), the float conversion points and the rule
"ZeroDivisionError / ValueError -> 1" (:535-543) are kept, because they decide the value in the
last bit and on degenerate tables.  Here the methods are a table of expressions over a `Table`;
they are evaluated with Python scalars (arbitrary-precision ints, IEEE doubles), so results equal
the reference's bit for bit -- including the complex numbers its `** .5` of a negative produces.
Pinned by tests/golden/cnidaria/ (oracle/make_golden_cnidaria.py runs the reference itself).

The N x N x 3 matrix of a .kma holds (Total_k, Total_l, Shared_kl) per pair: `apply` evaluates
methods over all pairs (count = total there).  The matrices are at most 255 x 255, host work.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np

_FIELDS = ("a,b,B,c,C,d,D,aPb,aPB,aPc,aPC,aPd,aPD,bPc,BPC,bPd,BPD,cPd,CPD,"
           "aTb,aTB,aTc,aTC,aTd,aTD,bTc,BTC,bTd,BTD,cTd,CTD,aPbPcPd,n,N,nTa").split(",")


class Table:
    """The 2x2 table of one pair and the sums / products the methods use
    (calculate_distances_cnidaria.py:491-532).  Sums are ints, products floats."""
    __slots__ = _FIELDS

    def __init__(self, totalX: int = 0, totalY: int = 0, countX: int = 0, countY: int = 0, val: int = 0):
        self.fill(totalX, totalY, countX, countY, val)

    def fill(self, totalX: int, totalY: int, countX: int, countY: int, val: int) -> "Table":
        a = self.a = val
        b = self.b = countX - val
        c = self.c = countY - val
        B = self.B = totalX - val
        C = self.C = totalY - val
        d = self.d = a + b + c                     # sic: not "neither", :499
        D = self.D = a + B + C
        for lhs, rhs in (("a", "b"), ("a", "B"), ("a", "c"), ("a", "C"), ("a", "d"), ("a", "D"),
                         ("b", "c"), ("B", "C"), ("b", "d"), ("B", "D"), ("c", "d"), ("C", "D")):
            x, y = getattr(self, lhs), getattr(self, rhs)
            setattr(self, f"{lhs}P{rhs}", x + y)
            setattr(self, f"{lhs}T{rhs}", float(x) * y)
        self.aPbPcPd = a + b + c + d
        self.n = a + b + c + d
        self.N = a + B + C + D
        self.nTa = float(self.n) * a
        return self

    def print_stats(self) -> None:
        for key in _FIELDS:
            print("%9s: %s" % (key, str(getattr(self, key))))


sqrt, log = math.sqrt, math.log


def _chi2(t: Table) -> float:          # pearson_chi_squared, :306-310
    return float(t.n * ((t.aTd - t.bTc) ** 2)) / (t.aPb * t.aPc * t.cPd * t.bPd)


def _phi(t: Table) -> float:           # pearson_phi, :312-316
    return (t.aTd - t.bTc) / sqrt(t.aPb * t.aPc * t.bPd * t.cPd)


def _sigma(t: Table):                  # :415-416
    return max(t.a, t.b) + max(t.c, t.d) + max(t.a, t.c) + max(t.b, t.d)


def _sigma_prime(t: Table):            # :418-419
    return max(t.aPc, t.bPd) + max(t.aPb, t.cPd)


def _s_jaccard(t: Table) -> float:
    return float(t.a) / (t.a + t.b + t.c)


def _hellinger_u(t: Table) -> float:
    return float(t.a) / sqrt(t.aPb * t.aPc)


def _cole(t: Table) -> float:
    diff = t.aTd - t.bTc
    return float(sqrt(2) * diff) / sqrt(diff ** 2 - (t.aPb * t.aPc * t.bPd * t.cPd))


def _sokal_sneath_iv(t: Table, third) -> float:
    parts = float(t.a) / t.aPb, float(t.a) / t.aPc, float(t.a) / third, float(t.a) / t.bPd
    return float(parts[0] + parts[1] + parts[2] + parts[3]) / 4


def _bub(t: Table, minus) -> float:
    root = sqrt(t.aTb)
    return float(root + t.a - minus) / (root + t.a + t.b + t.c)


# name -> (expression over the table, reported as "1 - value"?)
# Similarities (S_*) are reported as 1 - s; S_jaccard, S_intersection and S_innerproduct are
# reported raw, as are all D_* except D_yuleq -- exactly as in the reference.
_RAW, _ONE_MINUS = False, True
_METHODS: Dict[str, tuple] = {
    "S_jaccard": (_s_jaccard, _RAW),
    "D_jaccard": (lambda t: 1 - _s_jaccard(t), _RAW),
    "D_jaccard_sqrt": (lambda t: sqrt(1 - _s_jaccard(t)), _RAW),
    "S_dice": (lambda t: float(2.0 * t.a) / ((2.0 * t.a) + t.b + t.c), _ONE_MINUS),
    "S_jaccard3w": (lambda t: float(3.0 * t.a) / ((3.0 * t.a) + t.b + t.c), _ONE_MINUS),
    "S_nei_li": (lambda t: float(2.0 * t.a) / (t.aPb + t.aPc), _ONE_MINUS),
    "S_sokal_sneath_I": (lambda t: float(t.a) / (t.a + (2.0 * t.b) + (2.0 * t.c)), _ONE_MINUS),
    "S_sokal_michener": (lambda t: float(t.aPd) / t.aPbPcPd, _ONE_MINUS),
    "S_sokal_sneath_II": (lambda t: float(2.0 * t.aPd) / ((2.0 * t.a) + t.b + t.c + (2.0 * t.d)), _ONE_MINUS),
    "S_roger_tanimoto": (lambda t: float(t.aPd) / (t.a + (2.0 * (t.b + t.c)) + t.d), _ONE_MINUS),
    "S_faith": (lambda t: float(t.a + (0.5 * t.d)) / t.aPbPcPd, _ONE_MINUS),
    "S_gower_legendre": (lambda t: float(t.aPd) / (t.a + (0.5 * (t.b + t.c)) + t.d), _ONE_MINUS),
    "S_intersection": (lambda t: float(t.a), _RAW),
    "S_innerproduct": (lambda t: float(t.aPd), _RAW),
    "S_russell_rao": (lambda t: float(t.a) / t.aPbPcPd, _ONE_MINUS),
    "D_hamming": (lambda t: t.bPc, _RAW),
    "D_euclid": (lambda t: sqrt(t.bPc), _RAW),
    "D_squared_euclid": (lambda t: sqrt(t.bPc ** 2), _RAW),
    "D_mean_manhattan": (lambda t: float(t.bPc) / t.aPbPcPd, _RAW),
    "D_vari": (lambda t: float(t.bPc) / (4.0 * t.aPbPcPd), _RAW),
    "D_sized_difference": (lambda t: float(t.bPc ** 2) / (t.aPbPcPd ** 2), _RAW),
    "D_shaped_difference": (lambda t: float((t.n * t.bPc) - ((t.b - t.c) ** 2)) / (t.aPbPcPd ** 2), _RAW),
    "D_pattern_difference": (lambda t: float(4 * t.bTc) / (t.aPbPcPd ** 2), _RAW),
    "D_lance_williams": (lambda t: float(t.bPc) / ((2.0 * t.a) + t.b + t.c), _RAW),
    "D_bray_curtis": (lambda t: float(t.bPc) / ((2.0 * t.a) + t.b + t.c), _RAW),
    "D_hellinger": (lambda t: 2.0 * sqrt(1 - _hellinger_u(t)), _RAW),
    "D_chord": (lambda t: sqrt(2.0 * (1 - _hellinger_u(t))), _RAW),
    "S_cosine": (lambda t: float(t.a) / (sqrt(t.aPb * t.aPc) ** 2.0), _ONE_MINUS),
    "S_gilbert_wells": (lambda t: log(t.a) - log(t.n) - log(float(t.aPb) / t.n) - log(float(t.aPc) / t.n),
                        _ONE_MINUS),
    "S_ochiai_I": (_hellinger_u, _ONE_MINUS),
    "S_forbes_I": (lambda t: float(t.n * t.a) / (t.aPb * t.aPc), _ONE_MINUS),
    "S_fossum": (lambda t: float(t.n * ((t.a - .5) ** 2)) / (t.aPb * t.aPc), _ONE_MINUS),
    "S_sorgenfrei": (lambda t: float(t.a ** 2) / (t.aPb * t.aPc), _ONE_MINUS),
    "S_mountford": (lambda t: float(t.a) / ((0.5 * (t.aTb + t.aTc)) + t.bTc), _ONE_MINUS),
    "S_otsuka": (lambda t: float(t.a) / ((t.aPb * t.aPc) ** .5), _ONE_MINUS),
    "S_mcconnaughey": (lambda t: float((t.a ** 2) - t.bTc) / (t.aPb * t.aPc), _ONE_MINUS),
    "S_tarwid": (lambda t: float(t.nTa - (t.aPb * t.aPc)) / (t.nTa + (t.aPb * t.aPc)), _ONE_MINUS),
    "S_kulczynski_II": (lambda t: float((float(t.a) / 2) * ((2 * t.a) + t.b + t.c)) / (t.aPb * t.aPc), _ONE_MINUS),
    "S_driver_kroeber": (lambda t: (float(t.a) / 2.0) * ((1.0 / t.aPb) + (1.0 / t.aPc)), _ONE_MINUS),
    "S_johson": (lambda t: (float(t.a) / t.aPb) + (float(t.a) / t.aPc), _ONE_MINUS),
    "S_dennis": (lambda t: float(t.aTd - t.bTc) / sqrt(t.n * t.aPb * t.aPc), _ONE_MINUS),
    "S_simpson": (lambda t: float(t.a) / min([t.aPb, t.aPc]), _ONE_MINUS),
    "S_braun_banquet": (lambda t: float(t.a) / max([t.aPb, t.aPc]), _ONE_MINUS),
    "S_fager_mcgowan": (lambda t: _hellinger_u(t) - (max([t.aPb, t.aPc]) / 2.0), _ONE_MINUS),
    "S_forbes_II": (lambda t: float(t.nTa - (t.aPb * t.aPc)) / (t.n * min([t.aPb, t.aPc]) - (t.aPb * t.aPc)),
                    _ONE_MINUS),
    "S_sokal_sneath_IV": (lambda t: _sokal_sneath_iv(t, t.bPc), _ONE_MINUS),
    "S_sokal_sneath_IV2": (lambda t: _sokal_sneath_iv(t, t.bPd), _ONE_MINUS),
    "S_gower": (lambda t: float(t.aPd) / sqrt(t.aPb * t.aPc * t.bPd * t.cPd), _ONE_MINUS),
    "S_pearson_I": (_chi2, _ONE_MINUS),
    "S_pearson_II": (lambda t: (_chi2(t) / (t.n + _chi2(t))) ** .5, _ONE_MINUS),
    "S_pearson_III": (lambda t: (_phi(t) / (t.n + _phi(t))) ** .5, _ONE_MINUS),
    "S_pearson_heron_I": (_phi, _ONE_MINUS),
    "S_pearson_heron_II": (lambda t: math.cos(float(math.pi * sqrt(t.bTc)) / (sqrt(t.aTd) + sqrt(t.bTc))),
                           _ONE_MINUS),
    "S_sokal_sneath_III": (lambda t: float(t.aPd) / t.bPc, _ONE_MINUS),
    "S_sokal_sneath_V": (lambda t: float(t.aTd) / (t.aPb * t.aPc * t.bPd * (t.cPd ** .5)), _ONE_MINUS),
    "S_cole": (_cole, _ONE_MINUS),
    "S_ochiai_II": (lambda t: float(t.aTd) / sqrt(t.aPb * t.aPc * t.bPd * t.cPd), _ONE_MINUS),
    "S_yuleq": (lambda t: float(t.aTd - t.bTc) / float(t.aTd + t.bTc), _ONE_MINUS),
    "D_yuleq": (lambda t: (2.0 * t.bTc) / (t.aTd + t.bTc), _ONE_MINUS),
    "S_yulew": (lambda t: float(sqrt(t.aTd) - sqrt(t.bTc)) / float(sqrt(t.aTd) + sqrt(t.bTc)), _ONE_MINUS),
    "S_kulczynski_I": (lambda t: t.a / t.bPc, _ONE_MINUS),
    "S_tanimoto": (lambda t: t.a / (t.aPb + t.aPc - t.a), _ONE_MINUS),
    "S_dispersion": (lambda t: float(t.aTd - t.bTc) / ((t.a + t.b + t.c + t.d) ** 2), _ONE_MINUS),
    "S_hamann": (lambda t: float(t.aPd - t.bPc) / (t.a + t.b + t.c + t.d), _ONE_MINUS),
    "S_michael": (lambda t: float(4.0 * (t.aTd - t.bTc)) / ((t.aPb ** 2) + (t.bPc ** 2)), _ONE_MINUS),
    "S_goodman_kruskal": (lambda t: (_sigma(t) - _sigma_prime(t)) / ((2.0 * t.n) - _sigma_prime(t)), _ONE_MINUS),
    "S_anderberg": (lambda t: (_sigma(t) - _sigma_prime(t)) / (2.0 * t.n), _ONE_MINUS),
    "S_baroni_urbani_buser_I": (lambda t: _bub(t, 0), _ONE_MINUS),
    "S_baroni_urbani_buser_II": (lambda t: _bub(t, t.bPc), _ONE_MINUS),
    "S_pierce": (lambda t: float(t.aTb + t.bTc) / (t.aTb + (2 * t.bTc) + t.cTd), _ONE_MINUS),
    "S_eyraud": (lambda t: float((t.n ** 2) * (t.nTa - (t.aPb * t.aPc))) / (t.aPb * t.aPc * t.bPd * t.cPd),
                 _ONE_MINUS),
}


def _bind(name: str, table: Table) -> Callable[[], object]:
    expr, one_minus = _METHODS[name]
    if one_minus:
        return lambda: 1.0 - expr(table)
    return lambda: expr(table)


# the reference's module-level state: one shared table, every method bound to it
stats_data = Table()
methods_available: Dict[str, Callable] = {name: _bind(name, stats_data) for name in sorted(_METHODS)}
methods_enabled: Dict[str, Callable] = {}


def evaluate(name: str, totalX: int, totalY: int, countX: int, countY: int, val: int,
             table: Optional[Table] = None):
    """One method on one pair; degenerate tables give 1 (:535-543)."""
    t = (table or Table()).fill(totalX, totalY, countX, countY, val)
    expr, one_minus = _METHODS[name]
    try:
        r = expr(t)
        return 1.0 - r if one_minus else r
    except (ZeroDivisionError, ValueError):
        return 1


def attachMethodName(methodName: str, func: Callable) -> Callable:
    """-> f(dissi, num_kmers, x, y, totalX, totalY, countX, countY, val): dissi[methodName][x][y] += r
    (calculate_distances_cnidaria.py:475-549).  `func` reads the shared `stats_data`."""
    if methodName not in methods_enabled:
        print("unknown method:", methodName)
        raise KeyError

    def ffunc(dissi, num_kmers, x, y, totalX, totalY, countX, countY, val):
        stats_data.fill(totalX, totalY, countX, countY, val)
        try:
            r = func()
        except (ZeroDivisionError, ValueError):
            r = 1
        dissi[methodName][x][y] += r

    return ffunc


def init(methods_to_apply: Iterable[str]) -> None:
    """Enable methods by name; KeyError on an unknown one (:552-569)."""
    for m in methods_to_apply:
        if m not in methods_available:
            print(" unkknown method %s" % m)
            raise KeyError
        methods_enabled[m] = methods_available[m]
    for name in list(methods_enabled.keys()):
        methods_enabled[name] = attachMethodName(name, methods_available[name])


def apply(matrix: np.ndarray, methods: Optional[Sequence[str]] = None, fill_diagonal: bool = True) -> Dict[str, np.ndarray]:
    """.kma matrix (N, N, 3) of (Total_k, Total_l, Shared_kl) -> {method: (N, N) array}.
    count = total for both samples; the diagonal (which the reference's merger leaves uninitialised
    and its consumer zeroes, calculate_distance.py:96-97) is 0 when fill_diagonal."""
    names: List[str] = sorted(_METHODS) if methods is None else list(methods)
    for name in names:
        if name not in _METHODS:
            raise KeyError(name)
    m = np.asarray(matrix)
    n = m.shape[0]
    out = {name: [[0.0] * n for _ in range(n)] for name in names}
    table = Table()
    for k in range(n):
        for l in range(n):
            if k == l and fill_diagonal:
                continue
            tot_k, tot_l, shared = (int(v) for v in m[k, l])
            table.fill(tot_k, tot_l, tot_k, tot_l, shared)
            for name in names:
                expr, one_minus = _METHODS[name]
                try:
                    r = expr(table)
                    r = 1.0 - r if one_minus else r
                except (ZeroDivisionError, ValueError):
                    r = 1
                out[name][k][l] = r
    result = {}
    for name in names:
        arr = np.array(out[name])
        result[name] = arr if np.iscomplexobj(arr) else arr.astype(np.float64)
    return result
