"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(pc_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol(native):
    lib = native.load()
    declared = _declared()
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.pc_version() == 6
    assert lib.pc_launch_count() >= 0


def test_struct_layout_matches_header(native):
    # sizes the C compiler gives the same structs (natural alignment on x86-64)
    assert ctypes.sizeof(native.pc_segment) == 16
    assert ctypes.sizeof(native.pc_copy2d) == 48
    assert ctypes.sizeof(native.pc_attn_params) == 168
    assert ctypes.sizeof(native.pc_linear_params) == 176 and ctypes.sizeof(native.pc_chain_op) == 256
    import subprocess, tempfile, textwrap
    src = textwrap.dedent('''
        #include <stdio.h>
        #include "promptcache_b200.h"
        int main(void){ printf("%zu %zu %zu %zu %zu\\n", sizeof(pc_segment), sizeof(pc_copy2d), sizeof(pc_attn_params),
                                 sizeof(pc_linear_params), sizeof(pc_chain_op)); return 0; }
    ''')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [16, 48, 168, 176, 256]


def test_argument_validation_needs_no_gpu(native):
    """Invalid arguments are rejected on the host before anything touches the device."""
    lib = native.load()
    seg = (native.pc_segment * 1)()
    seg[0].base, seg[0].len, seg[0].dst_row = 16, 4, 0
    assert lib.pc_splice_kv(seg, 1, None, 1, 1, 64, 8, 2, None) == -1
    assert b"null" in lib.pc_last_error()
    assert lib.pc_splice_kv(seg, 1, 4096, 1, 1, 60, 8, 1, None) == -1          # row bytes not a multiple of 16
    assert lib.pc_splice_kv(seg, 1, 4096, 1, 1, 64, 3, 2, None) == -1          # rows exceed max_ctx
    assert b"max_ctx" in lib.pc_last_error()
    assert lib.pc_splice_kv(seg, 0, None, 1, 1, 64, 8, 2, None) == 0           # empty splice is a no-op
    p = native.pc_attn_params()
    assert lib.pc_prefill_attn(ctypes.byref(p), None) == 0                      # q_len == 0: nothing to do
    p.q_len = 2
    assert lib.pc_decode_attn(ctypes.byref(p), None) == -1                      # decode is q_len == 1
    ops = (native.pc_chain_op * 2)()
    assert lib.pc_linear_chain(ops, 0, None, 0, None) == -1                     # an empty chain
    assert lib.pc_linear_chain(ops, 9, None, 0, None) == -1                     # more than PC_CHAIN_MAX_OPS


def test_product_has_no_oracle_or_cpu_path():
    """Nothing under promptcache_b200/ may import oracle/, and CPU tensors are refused loudly."""
    import torch
    from promptcache_b200 import _native as nat
    for path in glob.glob(os.path.join(ROOT, "promptcache_b200", "**", "*.py"), recursive=True):
        text = open(path).read()
        assert "import oracle" not in text and "from oracle" not in text, path
    import pytest
    with pytest.raises(nat.NativeError):
        nat.splice_kv([], torch.zeros(1, 2, 1, 4, 64, dtype=torch.half))
