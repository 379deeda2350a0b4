"""Dependency probes of the drop-in ("Dependency probes to neutralise").

The reference's ``check_dependencies(args, runlogFile)`` (mirge/libs/miRgeEssential.py:6-98) exits unless the
``bowtie`` and ``cutadapt`` executables answer ``--version``, and it is the place where ``args.cutadaptVersion`` /
``args.bowtieVersion`` get their values (read by digest.py:111-113 and by the tRF / novel-miRNA modules).  With the B200
path neither tool runs, so:

* ``check_dependencies`` here has the same signature and side effects (prints, run.log lines, the two attributes)
  but probes what this path needs instead: the native library and a CUDA device;
* ``summarize()`` and ``bamFmt`` still shell out to ``bowtie-inspect`` (summary.py:776-788 ``-n``; :812-815 and
  :1164-1166 ``-a 20000 -e``; bamFmt.py:10-12 ``-n``).  ``write_inspect_shim(dir)`` writes a ``bowtie-inspect``
  executable that answers those calls from the library's FASTA or, when only the bowtie index ships, from the index
  files themselves (ebwt.py); pointing ``args.bowtie_path`` at ``dir`` lets the unmodified reference code run.

No alignment is ever routed through these shims: there is no ``bowtie`` stand-in, the rounds run in annotate.cu."""
from __future__ import annotations

import os
import stat
import sys
from typing import List, Sequence, Tuple

from . import REPO_ROOT

# cutadapt generation whose semantics the trim kernels restate (2.x - 3.x objective); >= 3 also
# selects the reference's newer call convention where it still consults the attribute (digest.py:112)
CUTADAPT_SEMANTICS = (3, 7)
# bowtie generation whose -v / -n / --best --strata policies the annotation kernel restates
BOWTIE_SEMANTICS = "1.3.1"

FASTA_EXTS = (".fa", ".fasta", ".fa.gz", ".fasta.gz")


def check_dependencies(args, runlogFile) -> None:
    """Drop-in for ``miRgeEssential.check_dependencies``: sets ``args.bowtieVersion`` / ``args.cutadaptVersion`` and
    verifies that the CUDA library loads and a device is present; user-facing failure = message + ``exit()``, as in
    the reference."""
    quiet = bool(getattr(args, "quiet", False))
    lines: List[str] = []
    problem = None
    try:
        from . import abi

        abi.load_library()
        import torch

        if not torch.cuda.is_available():
            problem = "mirge_b200 error!: no CUDA device is visible (the B200 path has no CPU fallback)"
        else:
            lines.append("mirge_b200 device: %s" % torch.cuda.get_device_name(torch.cuda.current_device()))
    except Exception as e:  # noqa: BLE001 -- reported to the user like a missing tool
        problem = "mirge_b200 error!: %s" % e
    with open(str(runlogFile), "a+") as outlog:
        if problem:
            print(problem)
            outlog.write(problem + "\n")
            outlog.close()
            sys.exit()  # the reference's probe ends with exit() after its message (miRgeEssential.py:22, :41)
        args.bowtieVersion = "True"
        args.cutadaptVersion = CUTADAPT_SEMANTICS
        lines.append("bowtie version: %s (semantics restated on the GPU; no bowtie process is run)" % BOWTIE_SEMANTICS)
        lines.append("cutadapt version: %s (semantics restated on the GPU; no cutadapt module is imported)"
                     % ".".join(str(v) for v in CUTADAPT_SEMANTICS))
        for ln in lines:
            if not quiet:
                print(ln)
            outlog.write(ln + "\n")


def _find_fasta(index_base: str):
    d, name = os.path.split(index_base)
    for folder in (d, os.path.join(os.path.dirname(d), "fasta.Libs")):
        for ext in FASTA_EXTS:
            p = os.path.join(folder, name + ext)
            if os.path.exists(p):
                return p
    return None


def _read_fasta_full(path: str) -> Tuple[List[str], List[bytes]]:
    """Whole header lines (what ``bowtie-inspect -n`` prints) and sequences."""
    import gzip

    op = gzip.open if path.endswith(".gz") else open
    names: List[str] = []
    seqs: List[List[bytes]] = []
    with op(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                names.append(line[1:].decode("latin-1"))
                seqs.append([])
            elif names and line:
                seqs[-1].append(line)
    return names, [b"".join(s) for s in seqs]


def _require_index(index_base: str) -> None:
    for ext in (".1.ebwt", ".3.ebwt"):
        if not os.path.exists(index_base + ext):
            raise FileNotFoundError("library %s is missing: no FASTA of that name and no %s" % (index_base, index_base + ext))


def index_entries(index_base: str) -> Tuple[List[str], List[bytes]]:
    """(full name lines, sequences) of the library addressed as a bowtie index base name: from the FASTA of that base
    name (index.Libs/ or fasta.Libs/) when there is one, else decoded from ``<base>.{1,3,4}.ebwt``."""
    fa = _find_fasta(index_base)
    if fa is not None:
        return _read_fasta_full(fa)
    _require_index(index_base)
    from . import ebwt

    return ebwt.decode_full(index_base)


def index_names(index_base: str) -> List[str]:
    """``bowtie-inspect -n <index_base>``."""
    fa = _find_fasta(index_base)
    if fa is not None:
        return _read_fasta_full(fa)[0]
    _require_index(index_base)
    from . import ebwt

    return ebwt.decode_names(index_base)


def inspect_main(argv: Sequence[str], out=None) -> int:
    """The subset of the ``bowtie-inspect`` command line the reference uses: ``-n`` / ``--names``, ``-a`` /
    ``--across <int>`` (line width of the FASTA output, default 60), ``-e`` / ``--ebwt-ref`` (accepted: both ways
    of reconstructing give the same text), ``-v`` / ``--verbose`` (ignored)."""
    out = out or sys.stdout
    names_only = False
    across = 60
    index = None
    i = 0
    argv = list(argv)
    while i < len(argv):
        a = argv[i]
        if a in ("-n", "--names"):
            names_only = True
        elif a in ("-a", "--across"):
            i += 1
            if i >= len(argv):
                sys.stderr.write("bowtie-inspect: -a needs an integer\n")
                return 1
            across = int(argv[i])
        elif a in ("-e", "--ebwt-ref", "-v", "--verbose"):
            pass
        elif a.startswith("-"):
            sys.stderr.write("bowtie-inspect (mirge_b200 shim): unsupported option %s\n" % a)
            return 1
        else:
            index = a
        i += 1
    if index is None:
        sys.stderr.write("bowtie-inspect (mirge_b200 shim): no index name given\n")
        return 1
    try:
        if names_only:
            for n in index_names(index):
                out.write(n + "\n")
            return 0
        names, seqs = index_entries(index)
    except Exception as e:  # noqa: BLE001 -- a command-line tool reports and fails
        sys.stderr.write("bowtie-inspect (mirge_b200 shim): %s\n" % e)
        return 1
    for n, s in zip(names, seqs):
        out.write(">" + n + "\n")
        t = s.decode("latin-1")
        if across <= 0:
            out.write(t + "\n")
        else:
            for p in range(0, len(t), across):
                out.write(t[p : p + across] + "\n")
    return 0


SHIM = '''#!%(python)s
# bowtie-inspect stand-in written by mirge_b200.essential.write_inspect_shim: answers the calls of miRge3.0's
# summarize() / bamFmt (-n; -a <int> -e) from the library FASTA or the bowtie index files.  No alignment runs here.
import sys
sys.path.insert(0, %(root)r)
import mirge_b200  # noqa: F401
from mirge_b200 import essential
sys.exit(essential.inspect_main(sys.argv[1:]))
'''


def write_inspect_shim(directory: str) -> str:
    """Write an executable ``bowtie-inspect`` into ``directory`` and return its path; use the directory as
    ``args.bowtie_path`` (``-pbwt``) so that summary.py:776 / bamFmt.py:10 find it."""
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, "bowtie-inspect")
    with open(path, "w") as f:
        f.write(SHIM % {"python": sys.executable, "root": REPO_ROOT})
    os.chmod(path, os.stat(path).st_mode | stat.S_IXUSR | stat.S_IXGRP | stat.S_IXOTH)
    return path
