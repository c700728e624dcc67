"""oracle/captures.py -- TEST INFRASTRUCTURE.  Where the tests find the reference's example captures.

The reference's ten bundled captures (345 MB of int16 samples) are data, not source.  They travel to the GPU
box as xz files (oracle/_ref/examples_full/<name>.tbin.xz, 107 MB in total, written in the build container by
`oracle/make_golden.py --stage-only` from /root/reference/examples); everything the tests read is derived from
those on demand, on whichever box the tests run:

  oracle/_ref/examples/<name>.full.tbin   the whole capture, byte-identical to the reference's file
  oracle/_ref/examples/<name>.tbin        the capture the per-segment event fixtures (tests/golden/*.segments.json)
                                          were generated from: the whole file for the small ones, the first
                                          `rows` rows + the end-of-data marker for the big ones

Nothing here is used by the product path, and nothing under tests/ reads /root/reference.
"""
from __future__ import annotations

import lzma
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RT_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")
FULL_XZ = os.path.join(OUT, "examples_full")
DERIVED = os.path.join(OUT, "examples")

# name, directory, Makefile options (examples/*/Makefile), rows of the prefix fixture (None = whole file)
EXAMPLES = [
    ("Microdata_20blks", "9trk_NRZI", "-v -m -nrzi -hex -ascii", None),
    ("PLAGO_beginning", "9trk_NRZI", "-v -m -nrzi -ips=50 -deskew -ebcdic -linefeed", 700_000),
    ("1600bpi_ukn_6s", "9trk_PE", "-v -m -ntrks=9 -pe -bpi=1600 -ips=50 -order=01234576p -tap -ascii -linefeed", 600_000),
    ("LJS009_part1_39blks", "9trk_PE", "-v -m -ntrks=9 -pe -bpi=1600 -ips=50 -tap -ebcdic -linesize=137", 600_000),
    ("1kblks_43blks", "9trk_GCR", "-v -m -gcr -ips=50 -order=76543210p -zeros -correct -tap -ascii -linefeed", 600_000),
    ("sf93_8blks", "9trk_GCR", "-v -m -gcr -ips=50 -zeros -correct -tap -ascii -linefeed", 600_000),
    ("analog", "9trk_GCR", "-v -gcr -ips=125 -differentiate -zeros -tap -ascii", None),
    ("SRI_SDS_102715028_4secs", "7trk_NRZI", "-v -m -nrzi -ntrks=7 -order=543210p -tap -SDS -linesize=144", 700_000),
    ("tss_4secs", "7trk_NRZI", "-v -m -nrzi -ntrks=7 -tap", 700_000),
    ("132_pt1", "6trk_Whirlwind", "-whirlwind -v3 -fluxdir=auto -tap -deskew -octal2 -flexo", None),
]
BY_NAME = {e[0]: e for e in EXAMPLES}
_XZ_FILTERS = [{"id": lzma.FILTER_LZMA2, "preset": 1}]


def reference_path(name: str) -> str:
    return os.path.join(REF, "examples", BY_NAME[name][1], name + ".tbin")


def xz_path(name: str) -> str:
    return os.path.join(FULL_XZ, name + ".tbin.xz")


def stage_xz(name: str) -> str:
    """build container only: compress the reference's capture into oracle/_ref/examples_full/"""
    dst = xz_path(name)
    if os.path.exists(dst):
        return dst
    os.makedirs(FULL_XZ, exist_ok=True)
    with open(reference_path(name), "rb") as src, lzma.open(dst + ".tmp", "wb", format=lzma.FORMAT_XZ, filters=_XZ_FILTERS) as out:
        shutil.copyfileobj(src, out, 1 << 22)
    os.replace(dst + ".tmp", dst)
    return dst


def full_path(name: str) -> str | None:
    """the whole capture (decompressed on first use); None if it was never staged"""
    dst = os.path.join(DERIVED, name + ".full.tbin")
    if os.path.exists(dst):
        return dst
    os.makedirs(DERIVED, exist_ok=True)
    tmp = dst + f".tmp{os.getpid()}"
    if os.path.exists(reference_path(name)):
        shutil.copyfile(reference_path(name), tmp)
    elif os.path.exists(xz_path(name)):
        with lzma.open(xz_path(name), "rb") as src, open(tmp, "wb") as out:
            shutil.copyfileobj(src, out, 1 << 22)
    else:
        return None
    os.replace(tmp, dst)
    return dst


def staged_path(name: str) -> str | None:
    """the capture behind tests/golden/<name>*.segments.json (a prefix of the big captures)"""
    from readtape_b200 import tbin
    import numpy as np
    dst = os.path.join(DERIVED, name + ".tbin")
    if os.path.exists(dst):
        return dst
    full = full_path(name)
    if full is None:
        return None
    rows = BY_NAME[name][3]
    tmp = dst + f".tmp{os.getpid()}"
    if rows is None:
        shutil.copyfile(full, tmp)
    else:
        hdr, allrows = tbin.read_tbin(full)
        with open(full, "rb") as fh:
            head = fh.read(hdr.payload_offset)
        with open(tmp, "wb") as fh:
            fh.write(head)
            np.array(allrows[:rows]).tofile(fh)
            fh.write(np.array([tbin.END_MARK], dtype="<i2").tobytes())
    os.replace(tmp, dst)
    return dst
