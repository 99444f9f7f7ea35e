"""Audio decode helpers on the ingest side of the hot path (host I/O, not compute).
``read_sph`` restates dataset/sph.py:33-63 (NIST SPHERE, big-endian 16-bit PCM) with a
vectorised decode instead of the reference's per-byte python loop."""
import numpy as np


def read_sph(fname):
    with open(fname, "rb") as f:
        head = f.read(1024)
        if not head.startswith(b"NIST_1A"):
            raise ValueError("not a NIST SPHERE file: " + fname)
        hsize = int(head.split(b"\n")[1].strip())
        if hsize > 1024:
            head += f.read(hsize - 1024)
        fields = {}
        for line in head.split(b"\n")[2:]:
            parts = line.strip().split()
            if not parts or parts[0] == b"end_head":
                break
            if len(parts) >= 3:
                fields[parts[0].decode()] = parts[2].decode()
        f.seek(hsize)
        raw = f.read()
    nbytes = int(fields.get("sample_n_bytes", 2))
    if nbytes != 2:
        raise ValueError("only 16-bit SPHERE PCM is supported")
    order = fields.get("sample_byte_format", "10")
    dt = ">i2" if order == "10" else "<i2"
    data = np.frombuffer(raw[: len(raw) // 2 * 2], dtype=dt).astype(np.int16)
    return int(fields.get("sample_rate", 16000)), data
