"""The reference's OWN programs -- encode.c, decode.c, benchmark.c -- compiled unchanged against this repository's
include/nanorq.h + libnanorq_b200.so (oracle/Makefile -> oracle/_ref/bin/*_b200), next to the same sources linked
against the unmodified reference (*_ref).  Wire-format interoperability both ways (the `data.rq` stream of encode.c /
decode.c: u64 oti_common, u32 oti_scheme, then {u32 tag, T bytes} per symbol), and the reference's benchmark with
its own output==input assertion."""
import os
import subprocess
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
have_bins = all(os.path.exists(os.path.join(BIN, n + s)) for n in ("encode", "decode", "benchmark") for s in ("_ref", "_b200"))
pytestmark = pytest.mark.skipif(not have_bins, reason="oracle/_ref/bin not built (needs the reference sources at build time)")


def run(name, args, cwd, timeout=600):
    return subprocess.run([os.path.join(BIN, name)] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True,
                          timeout=timeout)


def make_file(path, nbytes, seed):
    np.random.default_rng(seed).integers(0, 256, nbytes, dtype=np.uint8).tofile(path)


def test_reference_programs_with_the_reference_library(tmp_path):
    """Sanity of the harness itself (CPU only): encode_ref -> data.rq -> decode_ref."""
    make_file(tmp_path / "in.bin", 300007, 1)
    r = run("encode_ref", ["in.bin", 1280], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    r = run("decode_ref", ["out.bin"], tmp_path)
    assert r.returncode == 0 and "failed" not in r.stdout, r.stdout + r.stderr
    assert (tmp_path / "out.bin").read_bytes() == (tmp_path / "in.bin").read_bytes()


@pytest.mark.gpu
@pytest.mark.parametrize("nbytes,T", [(300007, 1280), (5 * 1024 * 1024 + 13, 1280), (40000, 64), (20 * 1024 * 1024, 1000)])
@pytest.mark.parametrize("enc,dec", [("encode_b200", "decode_ref"), ("encode_ref", "decode_b200"), ("encode_b200", "decode_b200")])
def test_wire_format_interoperates_with_the_reference(tmp_path, nbytes, T, enc, dec):
    """6 % random loss + 5 repair symbols per block (encode.c:27-28), time-seeded: whichever library encodes, the
    other decodes the stream to the same bytes."""
    make_file(tmp_path / "in.bin", nbytes, nbytes % 97)
    r = run(enc, ["in.bin", T], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    r = run(dec, ["out.bin"], tmp_path)
    assert r.returncode == 0 and "failed" not in r.stdout, r.stdout + r.stderr
    assert (tmp_path / "out.bin").read_bytes() == (tmp_path / "in.bin").read_bytes()


@pytest.mark.gpu
@pytest.mark.parametrize("T,K,pct", [(64, 10, 0), (1280, 100, 5.0), (1280, 1024, 5.0), (1280, 4096, 5.0)])
def test_reference_benchmark_runs_on_this_library(tmp_path, T, K, pct):
    """benchmark.c unchanged: four timed phases over 256 MiB each, then assert(in[i] == out[i]).  A time-seeded loss
    pattern with zero overhead is singular now and then (the reference exits with 'decode of sbn 0 failed' too):
    up to three attempts."""
    for attempt in range(3):
        r = run("benchmark_b200", [T, K, pct], tmp_path, timeout=900)
        if r.returncode == 0:
            break
        assert "decode of sbn" in r.stderr, r.stdout + r.stderr
        time.sleep(1.1)  # the program seeds rand() with time(0): another second, another loss pattern
    assert r.returncode == 0, r.stdout + r.stderr
    cols = r.stdout.split()
    assert int(cols[0]) == K and len(cols) == 5 and all(float(c) > 0 for c in cols[1:]), r.stdout
    print("benchmark %d %d %.1f -> K, encode, precalc-encode, decode, decode-oh [Mibit/s]: %s" % (T, K, pct, r.stdout.strip()))
