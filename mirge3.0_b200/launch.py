"""Run miRge3.0's own, unchanged command line on the B200 path.

``python -m mirge_b200.launch <miRge3.0 arguments>`` (after ``import mirge_b200``; or ``launch.run(argv)``) calls the
reference's ``mirge.__main__:main`` -- its argument parser, file validation, pickling, ``summarize`` and writers stay
the installed reference code -- with the three modules of the per-read hot path substituted before the reference
imports them (route 1, done by the import system instead of by editing ``__main__.py``):

    mirge.libs.digest.baking                    -> mirge_b200.digest.baking            (digest.py:105)
    mirge.libs.manifoldAlign.bwtAlign           -> mirge_b200.manifoldAlign.bwtAlign   (manifoldAlign.py:68)
    mirge.libs.miRgeEssential.check_dependencies -> mirge_b200.essential.check_dependencies (miRgeEssential.py:6)

Because the substitution happens before import, the box needs neither cutadapt / dnaio / xopen (imported at the top of
the reference's digest.py) nor a bowtie binary.  ``bakingEC`` (``-mEC``, digestEC.py) is outside this path: asking
for it stops with a message.  ``check_dependencies`` also points ``args.bowtie_path`` at a ``bowtie-inspect`` shim when
no real one is available, so that the reference's ``summarize()`` finds the miRNA names (summary.py:776-788)."""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import types
from typing import Optional, Sequence


def _bakingEC_unsupported(*_a, **_k):
    sys.exit("mirge_b200: -mEC (miREC error correction, mirge/libs/digestEC.py) is not part of the B200 hot path; "
             "run without -mEC")


def _tool(args, path_attr: str, name: str):
    """the executable the reference would run for ``name`` (its ``-pbwt`` / ``-psam`` / ``-prf`` directory, else PATH)"""
    d = getattr(args, path_attr, None)
    if d:
        cand = os.path.join(str(d), name)
        return cand if os.access(cand, os.X_OK) else None
    return shutil.which(name)


def downstream_tools(args):
    """[(tool, option that needs it)] the reference will shell out to AFTER the hot path, for the options of this run:
    -nmir: bowtie / bowtie-build / samtools / RNAfold (novel_mir.py:224-225,318-323); -ai and -trf: bowtie
    (mirge2_tRF_a2i.py:1056,1291); -bam: samtools (bamFmt.py:116,174).  The B200 path replaces none of these."""
    need = []
    if getattr(args, "novel_miRNA", False):
        need += [("bowtie_path", "bowtie", "-nmir"), ("bowtie_path", "bowtie-build", "-nmir"), ("samtools_path", "samtools", "-nmir"),
                 ("RNAfold_path", "RNAfold", "-nmir")]
    if getattr(args, "AtoI", False):
        need.append(("bowtie_path", "bowtie", "-ai"))
    if getattr(args, "tRNA_frag", False):
        need.append(("bowtie_path", "bowtie", "-trf"))
    if getattr(args, "bam_out", False):
        need.append(("samtools_path", "samtools", "-bam"))
    return need


def _check_dependencies(args, runlogFile):
    from . import essential

    essential.check_dependencies(args, runlogFile)
    # the probes of the reference's check_dependencies that still matter: tools its downstream modules run.  Missing
    # ones stop the run here, before the digest, instead of in the middle of summarize().
    missing = sorted({"%s (needed by %s)" % (name, opt) for attr, name, opt in downstream_tools(args) if _tool(args, attr, name) is None})
    if missing:
        sys.exit("mirge_b200: the requested options run tools that were not found: " + ", ".join(missing) +
                 ". Install them or point -pbwt / -psam / -prf at them; the B200 path only replaces cutadapt and the "
                 "annotation rounds' bowtie calls.")
    have = _tool(args, "bowtie_path", "bowtie-inspect")
    if have is None:  # summarize() / bamFmt shell out to it (summary.py:776, bamFmt.py:10)
        shim_dir = os.path.join(os.path.dirname(os.path.abspath(str(runlogFile))), ".mirge_b200_bin")
        essential.write_inspect_shim(shim_dir)
        # args.bowtie_path is also where the reference looks for bowtie / bowtie-build (novel_mir.py:318-321,
        # mirge2_tRF_a2i.py:1056): the shim directory must not hide real ones
        for name in ("bowtie", "bowtie-build"):
            real = _tool(args, "bowtie_path", name)
            link = os.path.join(shim_dir, name)
            if real is not None and not os.path.lexists(link):
                os.symlink(os.path.realpath(real), link)
        args.bowtie_path = shim_dir


def install() -> None:
    """Register the substitutes in ``sys.modules``.  Must run before ``mirge.__main__`` (or any reference module that
    imports the three modules) is imported; idempotent."""
    from . import digest, manifoldAlign

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__mirge_b200__ = True
        sys.modules[name] = m
        return m

    mod("mirge.libs.digest", baking=digest.baking)
    mod("mirge.libs.manifoldAlign", bwtAlign=manifoldAlign.bwtAlign)
    mod("mirge.libs.digestEC", bakingEC=_bakingEC_unsupported)
    # miRgeEssential also holds validate_files and the UID tables the rest of the reference uses: keep the reference's
    # module and replace the probe.  Its only third-party import is ``import cutadapt as ca`` (for ca.__version__).
    if "cutadapt" not in sys.modules:
        try:
            importlib.import_module("cutadapt")
        except Exception:  # noqa: BLE001 -- not installed: the module is only asked for its version string
            from . import essential

            mod("cutadapt", __version__=".".join(str(v) for v in essential.CUTADAPT_SEMANTICS))
    ess = importlib.import_module("mirge.libs.miRgeEssential")
    ess.check_dependencies = _check_dependencies


def run(argv: Optional[Sequence[str]] = None) -> None:
    """``miRge3.0 <argv>`` with the B200 hot path."""
    install()
    main_mod = importlib.import_module("mirge.__main__")
    main_mod.check_dependencies = _check_dependencies  # the name main() looks up (from-import made a copy)
    if argv is not None:
        sys.argv = ["miRge3.0"] + list(argv)
    main_mod.main()


if __name__ == "__main__":
    run()
