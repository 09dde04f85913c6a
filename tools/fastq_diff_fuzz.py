"""Differential fuzz of the native FASTQ loader (mb_fastq_load, host side of libmonica_b200.so) against the Python mirror of
Bio.SeqIO's reader (monica_b200/fastx.py): random small FASTQ files (CRLF, multi-line records, blank-padded and tabbed titles,
quality lines starting with '@' / '+'), then truncated / a line deleted / a byte flipped.  Both must give the same records or
both must refuse the file.  CPU only:  python tools/fastq_diff_fuzz.py [seed] [iterations]
Known, accepted differences: titles that START with white space (the native id is empty, SeqIO takes the first word) and
bare carriage returns inside a line."""
import sys, os, random, tempfile, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from monica_b200 import _lib, fastx
L = _lib.lib()
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
def native(path):
    fq = C.c_void_p()
    rc = L.mb_fastq_load(path.encode(), C.byref(fq))
    if rc != 0:
        return ("err", L.mb_last_error().decode(errors="replace")[:50])
    n = L.mb_fastq_n(fq)
    offp = C.POINTER(C.c_int64)()
    catp = L.mb_fastq_seqs(fq, C.byref(offp))
    off = np.ctypeslib.as_array(offp, shape=(n + 1,))
    cat = np.ctypeslib.as_array(catp, shape=(int(off[-1]),)).tobytes() if n and off[-1] else b""
    recs = []
    for i in range(n):
        ln, il = C.c_int64(), C.c_int32()
        hp = L.mb_fastq_header(fq, i, C.byref(ln), C.byref(il))
        h = C.string_at(hp, ln.value)
        recs.append((h[:il.value], h, cat[off[i]:off[i+1]]))
    L.mb_fastq_free(fq)
    return ("ok", recs)
def mirror(path):
    try:
        with open(path, "r", encoding="latin-1", newline="") as fh:
            return ("ok", [(r.id.encode("latin-1"), r.description.encode("latin-1"), str(r.seq).encode("latin-1")) for r in fastx.parse(fh, "fastq")])
    except ValueError as e:
        return ("err", str(e)[:50])
def base_file():
    out = []
    for i in range(random.randint(0, 6)):
        n = random.randint(0, 30)
        s = "".join(random.choice("ACGTN") for _ in range(n))
        t = f"r{i}" + random.choice(["", " c", "\tc d", "  ", " x\t"])
        nl = random.choice(["\n", "\n", "\r\n"])
        if random.random() < 0.2 and n > 4:
            k = n // 2
            out.append(f"@{t}{nl}{s[:k]}{nl}{s[k:]}{nl}+{random.choice(['', t])}{nl}{'I'*k}{nl}{'J'*(n-k)}{nl}")
        else:
            out.append(f"@{t}{nl}{s}{nl}+{nl}{''.join(random.choice('I@+J') for _ in range(n))}{nl}")
        if random.random() < 0.15: out.append(nl)
    return "".join(out).encode()
bad = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3000):
    b = bytearray(base_file())
    m = random.random()
    if m < 0.3 and b:        # truncate
        b = b[:random.randint(0, len(b))]
    elif m < 0.5 and b:      # delete a line
        lines = bytes(b).split(b"\n"); del lines[random.randrange(len(lines))]; b = bytearray(b"\n".join(lines))
    elif m < 0.6 and b:      # flip a byte
        b[random.randrange(len(b))] = random.choice(b"@+\nA \t")
    p = os.path.join(tempfile.gettempdir(), "mb_fastq_fuzz_%d.fastq" % os.getpid()); open(p, "wb").write(bytes(b))
    a, c = native(p), mirror(p)
    if a[0] != c[0] or (a[0] == "ok" and a[1] != c[1]):
        bad += 1
        if bad <= 8:
            print("DIFF", bytes(b)[:120], "\n  native:", a if a[0]=="err" else a[1][:3], "\n  mirror:", c if c[0]=="err" else c[1][:3])
print("iterations done, divergences:", bad)
