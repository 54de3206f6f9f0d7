"""The C ABI driven by a compiled host that holds arkworks-layout structs (tests/abi_host/driver.cpp: u64 Montgomery
limbs, Projective { x, y, z }, Affine { x, y, infinity }), i.e. the calls the Rust shim tools/ripp-b200 makes, without
Python, ctypes or torch in the process.  Inputs come from the oracle, outputs are checked against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import bls12_381 as E
from oracle import protocols as O
from oracle import synth as OS
from ripp_b200 import codec as C
from ripp_b200.ip_proofs import vk_enc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
    exe = os.path.join(tmp, "abi_driver")
    lib_dir = os.path.join(ROOT, "ripp_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "abi_host", "driver.cpp"), "-L", lib_dir, "-lripp_b200",
                    "-Wl,-rpath," + lib_dir], check=True)
    return exe


def test_cpp_host_with_arkworks_layout(tmp_path):
    n, m = 5, 4
    g1, g2 = OS.g1_points("host-a", n), OS.g2_points("host-b", n)
    g1[3] = None  # the identity: Projective with z = 0
    s, t = OS.scalars("host-s", n), OS.scalars("host-t", n)
    s[0], s[1] = 0, E.R - 1
    buf = struct.pack("<Q", n)
    buf += b"".join(C.g1_jac_enc(p, z=7 + i).tobytes() for i, p in enumerate(g1))   # non-trivial Z
    buf += b"".join(C.g2_jac_enc(p).tobytes() for p in g2)
    buf += C.fr_vec_enc(s).tobytes() + C.fr_vec_enc(t).tobytes()
    alpha, beta = OS.scalar("srs-alpha", 0), OS.scalar("srs-beta", 0)
    vk, proofs, inputs = OS.groth16_instance(m)
    buf += struct.pack("<Q", m) + C.fr_enc(alpha).tobytes() + C.fr_enc(beta).tobytes()
    for a, b, c in proofs:
        buf += C.g1_enc(a).tobytes() + C.g2_enc(b).tobytes() + C.g1_enc(c).tobytes() + bytes([a is None, b is None, c is None])
    vkw = np.ascontiguousarray(vk_enc(vk))
    buf += struct.pack("<QQ", len(inputs[0]), vkw.nbytes // 8) + vkw.tobytes()
    buf += b"".join(C.fr_vec_enc(row).tobytes() for row in inputs)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    fin.write_bytes(buf)
    exe = _build(str(tmp_path))
    r = subprocess.run([exe, str(fin), str(fout)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    out = fout.read_bytes()
    w = np.frombuffer(out[: 8 * (72 + 18 + 36 + 4)], dtype=np.uint32)
    assert C.gt_dec(w[:144]) == O.PairingInnerProduct.inner_product(g1, g2)
    assert C.g1_jac_dec(w[144:180]) == E.msm(g1, s, E.g1_add, E.g1_mul)
    assert C.g2_jac_dec(w[180:252]) == E.msm(g2, s, E.g2_add, E.g2_mul)
    assert C.fr_dec(w[252:260]) == sum(x * y for x, y in zip(s, t)) % E.R
    off = 8 * 130
    (mismatch,) = struct.unpack_from("<q", out, off)
    assert mismatch == -1  # RIPP_ERR_LEN_MISMATCH: InnerProductError::MessageLengthInvalid
    (plen,) = struct.unpack_from("<Q", out, off + 8)
    proof = out[off + 16: off + 16 + plen]
    want = O.aggregate_proofs(O.tipa_setup(m, alpha, beta), proofs)
    assert proof == O.ser_aggregate_proof(want)
    assert out[off + 16 + plen] == 1  # verify_aggregate_proof accepted
