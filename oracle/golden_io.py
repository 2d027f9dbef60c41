"""Shared (de)serialisation for the golden fixtures under tests/golden/.

TEST INFRASTRUCTURE ONLY (see oracle/ref_port.py header).  Exact rationals are
stored as [numerator, denominator] integer pairs in lowest terms with a positive
denominator; bulky results are stored as a SHA-256 digest of the canonical text
``"p/q;p/q;..."`` so that the fixtures stay small.
"""
import gzip
import hashlib
import json
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def digest_pq(pairs):
    h = hashlib.sha256()
    h.update(";".join("%d/%d" % (p, q) for p, q in pairs).encode())
    return h.hexdigest()


def save(name, obj):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".json.gz")
    raw = json.dumps(obj, separators=(",", ":")).encode()
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(raw)
    return path


def load(name):
    path = os.path.join(GOLDEN_DIR, name + ".json.gz")
    with gzip.open(path, "rb") as f:
        return json.loads(f.read().decode())
