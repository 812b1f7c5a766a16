"""Synthetic nanopore-like MIDSV workloads for parity tests and ``bench.py``.

Test/bench infrastructure only: nothing on the product path imports this package.  The rows come from the C
generator in ``midsv_synth.c`` (counter-based RNG keyed by read index, so shards are reproducible); this module
holds the workload definitions and the glue that turns a generated batch into the reference's JSONL wire format
(``{"QNAME","RNAME","MIDSV","STRAND","PRESET"}``, reference ``preprocess/alignment/midsv_caller.py:266-272``).
"""

from __future__ import annotations

import ctypes
import json
import subprocess
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "midsv_synth.c"
_LIB = _HERE / "libdj_synth.so"

SEED = 20240601


class _Feature(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("pos", ctypes.c_int32), ("len", ctypes.c_int32)]


class _Params(ctypes.Structure):
    _fields_ = [
        ("sub_rate", ctypes.c_double),
        ("del_rate", ctypes.c_double),
        ("ins_rate", ctypes.c_double),
        ("ins_geom_p", ctypes.c_double),
        ("homopolymer_x", ctypes.c_double),
        ("trunc_prob", ctypes.c_double),
        ("trunc_min", ctypes.c_double),
        ("trunc_max", ctypes.c_double),
        ("plus_strand", ctypes.c_double),
    ]


SUB, DEL, INS, INV = 1, 2, 3, 4

#: error profiles (sub, del, ins) — SURVEY.md §8d: R10-like is the benchmark profile, R9 matches the real fixture
PROFILES = {"r10": (0.005, 0.008, 0.004), "r9": (0.020, 0.034, 0.013), "clean": (0.0, 0.0, 0.0)}


def build(force: bool = False) -> Path:
    """Compile the generator with gcc (+OpenMP) next to its source."""
    if force or not _LIB.exists() or _LIB.stat().st_mtime < _SRC.stat().st_mtime:
        subprocess.run(
            ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", str(_LIB), str(_SRC)],
            check=True,
        )
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(str(_LIB))
        lib.dj_synth_generate.restype = ctypes.c_int64
        lib.dj_synth_generate.argtypes = [
            ctypes.POINTER(_Params), ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_int32, ctypes.c_char_p, ctypes.POINTER(_Feature), ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p,
        ]
        lib.dj_synth_qnames.restype = None
        lib.dj_synth_qnames.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_void_p]
        _lib = lib
    return _lib


def random_reference(length: int, seed: int = SEED) -> str:
    """i.i.d. uniform ACGT reference."""
    rng = np.random.default_rng(seed)
    return "".join(np.array(list("ACGT"))[rng.integers(0, 4, length)])


@dataclass
class Population:
    """Mixture of alleles expressed against one FASTA allele (the reference the reads were mapped to)."""

    reference: str
    alleles: list[list[tuple[int, int, int]]]  # per true allele: [(kind, pos, len), ...]
    fractions: list[float]
    profile: str = "r10"
    trunc_prob: float = 0.01
    rname: str = "control"


@dataclass
class MidsvBatch:
    """Packed rows: ``text[row_off[r] : row_off[r] + row_len[r]]`` is read r's MIDSV string."""

    text: np.ndarray  # uint8, rows separated by '\n', 64 bytes of slack at the end
    row_off: np.ndarray  # int64[N]
    row_len: np.ndarray  # int32[N]
    strand: np.ndarray  # uint8[N]  ('+' / '-')
    allele: np.ndarray  # uint8[N]  generating allele (ground truth, not used by the path)
    n_loci: int
    rname: str
    read_index: np.ndarray = field(default=None)  # int64[N] global read index
    seed: int = SEED

    @property
    def n_reads(self) -> int:
        return int(self.row_off.shape[0])

    def qnames(self) -> np.ndarray:
        out = np.empty(self.n_reads, dtype="S36")
        if self.n_reads:
            first = int(self.read_index[0])
            stride = int(self.read_index[1] - self.read_index[0]) if self.n_reads > 1 else 1
            _load().dj_synth_qnames(self.seed, first, stride, self.n_reads, out.ctypes.data)
        return out

    def midsv(self, r: int) -> str:
        o, n = int(self.row_off[r]), int(self.row_len[r])
        return self.text[o : o + n].tobytes().decode("ascii")

    def records(self):
        names = self.qnames()
        for r in range(self.n_reads):
            yield {
                "QNAME": names[r].decode("ascii"),
                "RNAME": self.rname,
                "MIDSV": self.midsv(r),
                "STRAND": chr(self.strand[r]),
                "PRESET": "map-ont",
            }

    def write_jsonl(self, path: str | Path) -> Path:
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        with open(path, "w") as f:
            for rec in self.records():
                f.write(json.dumps(rec) + "\n")
        return path


def qnames(seed: int, first_read: int, stride: int, n_reads: int) -> np.ndarray:
    """QNAMEs of reads first_read, first_read + stride, ... of the stream with this seed (a function of the index)."""
    out = np.empty(n_reads, dtype="S36")
    if n_reads:
        _load().dj_synth_qnames(seed, first_read, stride, n_reads, out.ctypes.data)
    return out


def generate(pop: Population, n_reads: int, seed: int = SEED, first_read: int = 0, stride: int = 1,
             text_alloc=None) -> MidsvBatch:
    """``text_alloc(nbytes)`` may supply the uint8 buffer for the text (e.g. a view of pinned host memory)."""
    lib = _load()
    sub, dele, ins = PROFILES[pop.profile]
    params = _Params(sub, dele, ins, 0.6, 3.0, pop.trunc_prob, 0.05, 0.30, 0.5)
    feats = [f for allele in pop.alleles for f in allele]
    arr = (_Feature * max(1, len(feats)))(*[_Feature(*f) for f in feats])
    first = np.zeros(len(pop.alleles) + 1, dtype=np.int32)
    first[1:] = np.cumsum([len(a) for a in pop.alleles])
    cdf = np.cumsum(np.asarray(pop.fractions, dtype=np.float64))
    cdf /= cdf[-1]
    L = len(pop.reference)
    ref = pop.reference.encode("ascii")
    row_len = np.zeros(n_reads, dtype=np.int32)
    row_off = np.zeros(n_reads, dtype=np.int64)
    strand = np.zeros(n_reads, dtype=np.uint8)
    allele = np.zeros(n_reads, dtype=np.uint8)
    args = (ctypes.byref(params), seed, first_read, stride, n_reads, L, ref, arr, first.ctypes.data,
            cdf.ctypes.data, len(pop.alleles))
    total = lib.dj_synth_generate(*args, None, 0, row_off.ctypes.data, row_len.ctypes.data,
                                  strand.ctypes.data, allele.ctypes.data)
    text = np.zeros(total + 64, dtype=np.uint8) if text_alloc is None else text_alloc(total + 64)
    text[total:] = 10
    got = lib.dj_synth_generate(*args, text.ctypes.data, total, row_off.ctypes.data, row_len.ctypes.data,
                                strand.ctypes.data, allele.ctypes.data)
    if got != total:
        raise RuntimeError("synthetic generator: buffer size mismatch")
    idx = first_read + stride * np.arange(n_reads, dtype=np.int64)
    return MidsvBatch(text, row_off, row_len, strand, allele, L, pop.rname, idx, seed)


# ---------------------------------------------------------------------------------------------------------
# Workloads
# ---------------------------------------------------------------------------------------------------------


def throughput_population(length: int = 5000, profile: str = "r10", scale: float | None = None) -> Population:
    """8 alleles at 40/25/15/10/5/2/2/1 % inside one FASTA allele "control" (positions scale with length)."""
    s = (length / 5000.0) if scale is None else scale
    p = lambda x: max(1, min(length - 40, int(round(x * s))))  # noqa: E731
    alleles = [
        [],
        [(SUB, p(1000), 1)],
        [(DEL, p(1800), 1)],
        [(INS, p(2400), 1)],
        [(DEL, p(3000), 30)],
        [(INS, p(3500), 12)],
        [(SUB, p(600), 1), (SUB, p(4200), 1)],
        [(DEL, p(4600), 2)],
    ]
    return Population(random_reference(length), alleles, [40, 25, 15, 10, 5, 2, 2, 1], profile)


def control_population(length: int = 5000, profile: str = "r10", reference: str | None = None) -> Population:
    return Population(reference or random_reference(length), [[]], [1.0], profile)


def point_mutation_population(length: int, fraction: float, profile: str = "r10") -> Population:
    """batch-like: one substitution allele at the given fraction."""
    return Population(random_reference(length), [[], [(SUB, length // 2, 1)]], [1 - fraction, fraction], profile)


def inversion_population(length: int = 2724, profile: str = "r10") -> Population:
    """inversion-like: a ~600-nt lowercase block in every read."""
    return Population(random_reference(length), [[(INV, 1000, 600)]], [1.0], profile)
