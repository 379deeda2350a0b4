"""Decoder for bowtie 1 index files (``<base>.1.ebwt/.3.ebwt/.4.ebwt``): recovers (names, sequences)
so that a stock ``miRge3_Lib`` (which ships only bowtie indexes, manifoldAlign.py:97-98) can be loaded
without bowtie-inspect.  SURVEY.md section 8f row 1 -- scheduled after the hot path; not built yet."""
from .device import MirgeError


def decode_index(basename: str):
    raise MirgeError(
        "loading %s.*.ebwt is not implemented yet: put the library FASTA next to the index "
        "(<basename>.fa) -- see INTEGRATION.md" % basename
    )
