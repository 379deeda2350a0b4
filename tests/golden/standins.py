"""Stand-ins for the third-party packages the reference imports (cutadapt, dnaio, xopen, Bio), used ONLY by
make_reference_golden.py to drive the reference's own code in a container that has none of them installed.

Each stand-in implements just the API surface mirge/libs/digest.py touches, with the oracle's restatement of
the package's semantics (oracle/pyoracle.py: cutadapt's quality trimming and Aligner.locate, dnaio's chunking
and FASTQ parsing).  So the fixtures pin everything the REFERENCE does around those calls -- modifier order,
per-modifier (HEAD) counting, the qiagen split trick, UMIParser, the parent merge, both UMI levels, the
matrix build, the counters, the side files -- and do not pin the third-party arithmetic itself."""
import gzip
import io
import sys
import types

from mirge_b200 import params as P
from oracle import pyoracle as po


class Sequence:
    """dnaio.Sequence: name / sequence / qualities with slicing."""

    __slots__ = ("name", "sequence", "qualities")

    def __init__(self, name, sequence, qualities=None):
        self.name, self.sequence, self.qualities = name, sequence, qualities

    def __getitem__(self, key):
        return Sequence(self.name, self.sequence[key], None if self.qualities is None else self.qualities[key])

    def __len__(self):
        return len(self.sequence)


class _FastqReader:
    def __init__(self, fileobj):
        self.data = fileobj.read()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __iter__(self):
        for name, seq, qual in po.parse_fastq(self.data):
            yield Sequence(name, seq, qual)


def _read_chunks(f, buffer_size=4 * 1024 ** 2):
    data = f.read()
    mv = memoryview(data)
    for s, e in po.read_chunks(data, buffer_size):
        yield mv[s:e]


def _xopen(path, mode="rb", **kw):
    with open(path, "rb") as fh:
        magic = fh.read(2)
    return gzip.open(path, mode) if magic == b"\x1f\x8b" else open(path, mode)


class ModificationInfo:
    def __init__(self, read):
        self.matches = []


class NextseqQualityTrimmer:
    def __init__(self, cutoff, base):
        self.cutoff, self.base = cutoff, base

    def __call__(self, read, info=None):
        return read[: po.nextseq_trim_index(read.sequence, read.qualities, self.cutoff, self.base)]


class QualityTrimmer:
    def __init__(self, cutoff_front, cutoff_back, base):
        self.cf, self.cb, self.base = cutoff_front, cutoff_back, base

    def __call__(self, read, info=None):
        s, e = po.quality_trim_index(read.qualities, self.cf, self.cb, self.base)
        return read[s:e]


class AdapterCutter:
    def __init__(self, adapters, times=1, action="trim"):
        assert action == "trim"
        self.adapters, self.times = adapters, times

    def __call__(self, read, info=None):
        for _ in range(self.times):
            ad, mt = po.best_match(self.adapters, read.sequence)
            if mt is None:
                break
            if ad.where == "linked":
                read = read[mt[2] : mt[3]]
            else:
                read = read[mt[3]:] if ad.where in po.REMOVE_BEFORE else read[: mt[2]]
        return read


class NEndTrimmer:
    def __call__(self, read, info=None):
        s, e = 0, len(read.sequence)
        while s < e and read.sequence[s] == "N":
            s += 1
        while e > s and read.sequence[e - 1] == "N":
            e -= 1
        return read[s:e]


class UnconditionalCutter:
    def __init__(self, length):
        self.length = length

    def __call__(self, read, info=None):
        return read[self.length:] if self.length > 0 else read[: self.length]


def _make_adapters_from_specifications(specs, search_parameters):
    from tests.util import py_adapters

    return py_adapters(P.TrimConfig(adapters=list(specs), error_rate=search_parameters["max_errors"], overlap=search_parameters["min_overlap"],
                                    indels=search_parameters["indels"], match_adapter_wildcards=search_parameters["adapter_wildcards"],
                                    match_read_wildcards=search_parameters["read_wildcards"]))


def install(reference_root):
    """Register the stand-in modules and put the reference checkout on sys.path."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    unused = type("Unused", (), {})
    mod("cutadapt", __version__="3.1 (stand-in)")
    mod("cutadapt.adapters", warn_duplicate_adapters=lambda adapters: None)
    mod("cutadapt.modifiers", LengthTagModifier=unused, SuffixRemover=unused, PrefixSuffixAdder=unused, ZeroCapper=unused,
        QualityTrimmer=QualityTrimmer, UnconditionalCutter=UnconditionalCutter, NEndTrimmer=NEndTrimmer, AdapterCutter=AdapterCutter,
        PairedAdapterCutterError=unused, PairedAdapterCutter=unused, NextseqQualityTrimmer=NextseqQualityTrimmer, Shortener=unused,
        ModificationInfo=ModificationInfo)
    mod("cutadapt.parser", make_adapters_from_specifications=_make_adapters_from_specifications)
    mod("dnaio", read_chunks=_read_chunks, open=lambda f, **kw: _FastqReader(f), Sequence=Sequence)
    mod("xopen", xopen=_xopen)
    bio = mod("Bio")
    bio.Seq = mod("Bio.Seq", Seq=object)
    bio.SeqIO = mod("Bio.SeqIO")
    bio.pairwise2 = mod("Bio.pairwise2")
    if str(reference_root) not in sys.path:
        sys.path.insert(0, str(reference_root))
