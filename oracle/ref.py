"""Runs oracle/_ref/flexref (the reference's own host sources compiled for the CPU, see
ref_build.sh / ref_driver.cc) and parses its dumps.  TEST INFRASTRUCTURE ONLY."""
import os
import struct
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FLEXREF = os.path.join(_HERE, "_ref", "flexref")
_DT = {"I": np.uint32, "i": np.int32, "f": np.float32, "q": np.int64, "Q": np.uint64, "d": np.float64}


def available():
    return os.path.exists(FLEXREF)


def build():
    subprocess.call(["bash", os.path.join(_HERE, "ref_build.sh")])


def parse(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    o = 0
    while o < len(data):
        (l,) = struct.unpack_from("<I", data, o); o += 4
        name = data[o:o + l].decode(); o += l
        dt = chr(data[o]); o += 1
        (cnt,) = struct.unpack_from("<Q", data, o); o += 8
        a = np.frombuffer(data, _DT[dt], cnt, o).copy(); o += a.nbytes
        out[name] = a[0].item() if dt in "qd" and cnt == 1 else a
    return out


def run(cmd, csv, *args, timeout=600):
    """Returns the parsed dump, or raises RuntimeError (with the exit code) if the reference
    asserted/aborted on this input."""
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as t:
        outp = t.name
    try:
        r = subprocess.run([FLEXREF, cmd, csv, outp, *map(str, args)], capture_output=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"flexref {cmd} exited {r.returncode}: {r.stderr.decode()[-400:]}")
        return parse(outp)
    finally:
        if os.path.exists(outp):
            os.unlink(outp)


def write_csv(path, rowptr, col, val):
    """3-line CSV (DataLoader.cu:19-53); values printed with 9 significant digits (exact fp32)."""
    with open(path, "w") as f:
        f.write(",".join(map(str, np.asarray(rowptr).tolist())) + "\n")
        f.write(",".join(map(str, np.asarray(col).tolist())) + "\n")
        f.write(",".join("%.9g" % x for x in np.asarray(val, np.float32).tolist()) + "\n")
