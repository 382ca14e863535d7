#!/usr/bin/env python3
"""Pack the F-16 aero database into the binary/include forms the B200 library consumes.

Reads a checkout of the reference project (default /root/reference) and writes

  f16_mpc_oop_py_b200/data/f16_aero_v1.bin      hifi (Nguyen) breakpoints + 43 tables, FP64
  f16_mpc_oop_py_b200/csrc/f16_lofi_data.inc    lofi (Stevens-Lewis) tables as a flat FP64 array

Sources parsed (data only, no code is copied):
  C/ALPHA1.dat ALPHA2.dat BETA1.dat DH1.dat DH2.dat and the 43 C/*.dat tables that
  hifi_F16_AeroData.c opens (hifi_F16_AeroData.c:8,29,50,70,90,136,...,1848);
  the numeric initialisers of the arrays in lofi_F16_AeroData.c:17-26,66-104,192-206,271-283,343-344.

Blob format (little endian):
  8  bytes  magic  b"F16AERO1"
  8  bytes  uint64 number of doubles that follow (= 61 + 13405)
  32 bytes  sha256 of the payload bytes
  payload   doubles: ALPHA1[20] ALPHA2[14] BETA1[19] DH1[5] DH2[3] then the tables in the order
            of HIFI_TABLES below, each in the reference's own column-major order (alpha fastest).

The table order here must stay in sync with csrc/f16_tables.h (F16_CANON_*).
"""
import argparse
import hashlib
import os
import re
import struct
import sys

import numpy as np

BREAKPOINTS = [("ALPHA1", 20), ("ALPHA2", 14), ("BETA1", 19), ("DH1", 5), ("DH2", 3)]

# (canonical name, .dat file, shape (alpha, beta, dele)), reference accessor (hifi_F16_AeroData.c line of fopen)
HIFI_TABLES = [
    # A1 x B x D1
    ("Cx", "CX0120_ALPHA1_BETA1_DH1_201.dat", (20, 19, 5)),   # _Cx :136
    ("Cz", "CZ0120_ALPHA1_BETA1_DH1_301.dat", (20, 19, 5)),   # _Cz :183
    ("Cm", "CM0120_ALPHA1_BETA1_DH1_101.dat", (20, 19, 5)),   # _Cm :227
    # A1 x B x D2
    ("Cn", "CN0120_ALPHA1_BETA1_DH2_501.dat", (20, 19, 3)),   # _Cn :314
    ("Cl", "CL0120_ALPHA1_BETA1_DH2_601.dat", (20, 19, 3)),   # _Cl :359
    # A1 x B
    ("Cy", "CY0320_ALPHA1_BETA1_401.dat", (20, 19)),          # _Cy :270
    ("Cy_r30", "CY0720_ALPHA1_BETA1_405.dat", (20, 19)),      # :1356
    ("Cn_r30", "CN0720_ALPHA1_BETA1_503.dat", (20, 19)),      # :1398
    ("Cl_r30", "CL0720_ALPHA1_BETA1_603.dat", (20, 19)),      # :1440
    ("Cy_a20", "CY0620_ALPHA1_BETA1_403.dat", (20, 19)),      # :1482
    ("Cn_a20", "CN0620_ALPHA1_BETA1_504.dat", (20, 19)),      # :1566
    ("Cl_a20", "CL0620_ALPHA1_BETA1_604.dat", (20, 19)),      # :1650
    # A2 x B
    ("Cx_lef", "CX0820_ALPHA2_BETA1_202.dat", (14, 19)),      # :402
    ("Cz_lef", "CZ0820_ALPHA2_BETA1_302.dat", (14, 19)),      # :444
    ("Cm_lef", "CM0820_ALPHA2_BETA1_102.dat", (14, 19)),      # :486
    ("Cy_lef", "CY0820_ALPHA2_BETA1_402.dat", (14, 19)),      # :528
    ("Cn_lef", "CN0820_ALPHA2_BETA1_502.dat", (14, 19)),      # :570
    ("Cl_lef", "CL0820_ALPHA2_BETA1_602.dat", (14, 19)),      # :612
    ("Cy_a20_lef", "CY0920_ALPHA2_BETA1_404.dat", (14, 19)),  # :1524
    ("Cn_a20_lef", "CN0920_ALPHA2_BETA1_505.dat", (14, 19)),  # :1608
    ("Cl_a20_lef", "CL0920_ALPHA2_BETA1_605.dat", (14, 19)),  # :1691
    # A1
    ("CXq", "CX1120_ALPHA1_204.dat", (20,)),                  # :652
    ("CZq", "CZ1120_ALPHA1_304.dat", (20,)),                  # :691
    ("CMq", "CM1120_ALPHA1_104.dat", (20,)),                  # :730
    ("CYp", "CY1220_ALPHA1_408.dat", (20,)),                  # :769
    ("CYr", "CY1320_ALPHA1_406.dat", (20,)),                  # :808
    ("CNr", "CN1320_ALPHA1_506.dat", (20,)),                  # :847
    ("CNp", "CN1220_ALPHA1_508.dat", (20,)),                  # :886
    ("CLp", "CL1220_ALPHA1_608.dat", (20,)),                  # :925
    ("CLr", "CL1320_ALPHA1_606.dat", (20,)),                  # :964
    ("dCNbeta", "CN9999_ALPHA1_brett.dat", (20,)),            # :1731
    ("dCLbeta", "CL9999_ALPHA1_brett.dat", (20,)),            # :1770
    ("dCm", "CM9999_ALPHA1_brett.dat", (20,)),                # :1809
    # A2
    ("dCXq_lef", "CX1420_ALPHA2_205.dat", (14,)),             # :1003
    ("dCYr_lef", "CY1620_ALPHA2_407.dat", (14,)),             # :1042
    ("dCYp_lef", "CY1520_ALPHA2_409.dat", (14,)),             # :1081
    ("dCZq_lef", "CZ1420_ALPHA2_305.dat", (14,)),             # :1120
    ("dCLr_lef", "CL1620_ALPHA2_607.dat", (14,)),             # :1159
    ("dCLp_lef", "CL1520_ALPHA2_609.dat", (14,)),             # :1198
    ("dCMq_lef", "CM1420_ALPHA2_105.dat", (14,)),             # :1237
    ("dCNr_lef", "CN1620_ALPHA2_507.dat", (14,)),             # :1276
    ("dCNp_lef", "CN1520_ALPHA2_509.dat", (14,)),             # :1315
    # D1
    ("eta_el", "ETA_DH1_brett.dat", (5,)),                    # :1848
]

MAGIC = b"F16AERO1"

# lofi arrays in the order they are emitted; (function, name, rows) as they appear in lofi_F16_AeroData.c
LOFI_ARRAYS = [
    ("damping", "A", 9),     # :17   9 x 12
    ("dmomdcon", "ALA", 7),  # :66   7 x 12 (an 8th zero row is appended, see f16_tables.h)
    ("dmomdcon", "ALR", 7),  # :76
    ("dmomdcon", "ANA", 7),  # :86
    ("dmomdcon", "ANR", 7),  # :96
    ("clcn", "AL", 7),       # :192
    ("clcn", "AN", 7),       # :200
    ("cxcm", "AX", 5),       # :271
    ("cxcm", "AM", 5),       # :278
    ("cz", "A", 1),          # :343  12
]


def read_dat(path, n):
    vals = np.array(open(path).read().split(), dtype=np.float64)
    if vals.size != n:
        raise SystemExit(f"{path}: expected {n} values, found {vals.size}")
    return vals


def canonical_payload(cdir):
    parts = []
    for name, n in BREAKPOINTS:
        parts.append(read_dat(os.path.join(cdir, name + ".dat"), n))
    for _, fname, shape in HIFI_TABLES:
        parts.append(read_dat(os.path.join(cdir, fname), int(np.prod(shape))))
    return np.concatenate(parts)


def payload_size():
    return sum(n for _, n in BREAKPOINTS) + sum(int(np.prod(s)) for _, _, s in HIFI_TABLES)


def write_blob(payload, out):
    raw = payload.astype("<f8").tobytes()
    with open(out, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", payload.size))
        f.write(hashlib.sha256(raw).digest())
        f.write(raw)


def read_blob(path):
    b = open(path, "rb").read()
    if b[:8] != MAGIC:
        raise ValueError("bad magic in " + path)
    (n,) = struct.unpack("<Q", b[8:16])
    sha = b[16:48]
    raw = b[48:48 + 8 * n]
    if hashlib.sha256(raw).digest() != sha:
        raise ValueError("sha256 mismatch in " + path)
    return np.frombuffer(raw, dtype="<f8").copy()


def split_payload(payload):
    """-> (dict of breakpoints, dict of tables as Fortran-ordered arrays [alpha, beta, dele])."""
    off = 0
    bps, tabs = {}, {}
    for name, n in BREAKPOINTS:
        bps[name] = payload[off:off + n]
        off += n
    for name, _, shape in HIFI_TABLES:
        n = int(np.prod(shape))
        tabs[name] = payload[off:off + n].reshape(shape, order="F")
        off += n
    assert off == payload.size
    return bps, tabs


def parse_lofi_arrays(src_path):
    """Return the numeric initialisers of lofi_F16_AeroData.c keyed by (function, array name)."""
    src = open(src_path).read()
    out = {}
    # split into functions so that the two arrays both called 'A' (damping :17, cz :343) stay apart
    func_pat = re.compile(r"\nvoid\s+(\w+)\s*\(")
    starts = [(m.group(1), m.start()) for m in func_pat.finditer(src)]
    starts.append(("", len(src)))
    num = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?")
    arr = re.compile(r"double\s+(\w+)\s*(?:\[[^\]]*\])+\s*=\s*\{(.*?)\};", re.S)
    for (fname, a), (_, b) in zip(starts[:-1], starts[1:]):
        for m in arr.finditer(src[a:b]):
            out[(fname, m.group(1))] = np.array([float(t) for t in num.findall(m.group(2))])
    return out


def write_lofi_inc(cdir, out):
    arrays = parse_lofi_arrays(os.path.join(cdir, "lofi_F16_AeroData.c"))
    flat = []
    for func, name, rows in LOFI_ARRAYS:
        a = arrays[(func, name)]
        assert a.size == rows * 12, (func, name, a.size)
        flat.append(a)
        if func == "dmomdcon":
            flat.append(np.zeros(12))  # row 8: read with weight 0 at |beta| == 30 (lofi_F16_AeroData.c:136-150)
    flat = np.concatenate(flat)
    with open(out, "w") as f:
        f.write("// GENERATED by tools/pack_tables.py from the Stevens-Lewis lofi tables of the reference\n")
        f.write("// (numeric data only; layout documented in f16_tables.h, F16_LOFI_*). Do not edit.\n")
        f.write(f"// {flat.size} doubles, sha256(le f8) = {hashlib.sha256(flat.astype('<f8').tobytes()).hexdigest()}\n")
        for i in range(0, flat.size, 6):
            f.write(" ".join(f"{float(v)!r}," for v in flat[i:i + 6]) + "\n")
    return flat


def write_dat_dir(payload, outdir):
    """Regenerate C/*.dat text files (for the reference .so under oracle/_ref) from a payload."""
    os.makedirs(outdir, exist_ok=True)
    bps, tabs = split_payload(payload)
    for name, _ in BREAKPOINTS:
        with open(os.path.join(outdir, name + ".dat"), "w") as f:
            f.write(" ".join(repr(float(v)) for v in bps[name]) + "\n")
    for name, fname, _ in HIFI_TABLES:
        with open(os.path.join(outdir, fname), "w") as f:
            f.write(" ".join(repr(float(v)) for v in tabs[name].ravel(order="F")) + "\n")


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(here)
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(repo, "f16_mpc_oop_py_b200", "data", "f16_aero_v1.bin"))
    ap.add_argument("--lofi-inc", default=os.path.join(repo, "f16_mpc_oop_py_b200", "csrc", "f16_lofi_data.inc"))
    args = ap.parse_args()
    cdir = os.path.join(args.reference, "C")
    payload = canonical_payload(cdir)
    assert payload.size == payload_size() == 13466
    write_blob(payload, args.out)
    assert np.array_equal(read_blob(args.out), payload)
    lofi = write_lofi_inc(cdir, args.lofi_inc)
    print(f"wrote {args.out}: {payload.size} doubles; {args.lofi_inc}: {lofi.size} doubles")


if __name__ == "__main__":
    sys.exit(main())
