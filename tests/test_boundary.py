"""Drop-in boundary: the C-ABI library loads, exports exactly what include/prism_b200.h declares,
the ctypes mirror agrees with it, and the product never routes through the oracle."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "prism_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = re.findall(r"\b(?:int|long long|const char \*)\s*\**\s*(pb_[a-z0-9_]+)\s*\(([^;{]*)\)\s*;", src)
    return {name: [a.strip() for a in args.split(",")] if args.strip() not in ("", "void") else []
            for name, args in decls}


def test_library_exports_every_declared_symbol():
    from prism_b200 import _lib
    lib = _lib.load()
    fns = header_functions()
    assert len(fns) >= 25
    for name in fns:
        assert hasattr(lib, name), name
    assert lib.pb_abi_version() == 2
    assert lib.pb_error_string(-1).decode() == "invalid argument"


def test_ctypes_mirror_matches_header_arity():
    from prism_b200 import _lib
    fns = header_functions()
    assert set(fns) == set(_lib.SIGNATURES), set(fns) ^ set(_lib.SIGNATURES)
    for name, args in fns.items():
        assert len(args) == len(_lib.SIGNATURES[name]), name


def test_struct_layouts(tmp_path):
    """ctypes Structures vs the C compiler's view of include/prism_b200.h."""
    import subprocess
    from prism_b200 import _lib
    assert C.sizeof(_lib.pb_per_state) == 64
    assert _lib.pb_per_state.p_sum.offset == 20 and _lib.pb_per_state.p_min.offset == 24
    prog = tmp_path / "sz.c"
    prog.write_text(r"""
#include <stdio.h>
#include <stddef.h>
#include "prism_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(pb_per_state), sizeof(pb_tree), sizeof(pb_store),
         offsetof(pb_tree, eps_f64), offsetof(pb_tree, default_priority_fp64),
         offsetof(pb_store, size), offsetof(pb_store, gamma),
         sizeof(pb_peer_group), offsetof(pb_peer_group, reduced), offsetof(pb_peer_group, state),
         offsetof(pb_peer_group, epoch));
  return 0; }
""")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_lib.pb_per_state), C.sizeof(_lib.pb_tree), C.sizeof(_lib.pb_store),
            _lib.pb_tree.eps_f64.offset, _lib.pb_tree.default_priority_fp64.offset,
            _lib.pb_store.size.offset, _lib.pb_store.gamma.offset,
            C.sizeof(_lib.pb_peer_group), _lib.pb_peer_group.reduced.offset, _lib.pb_peer_group.state.offset,
            _lib.pb_peer_group.epoch.offset]
    assert got == want
    assert _lib.PB_PEER_MAX == 8


def test_peer_block_layout_and_slices():
    """Host-side arithmetic of the peer exchange: block layout (prism_b200/peer.py) and the rank slices
    (pb_peer_slice): 16-byte aligned, disjoint, covering."""
    from prism_b200 import _lib, peer
    lib = _lib.load()
    for n in (4, 4096, 265_220, 18_167_208):
        off = peer._layout(n)
        assert off["grad"] == 0 and off["reduced"] >= 4 * n and off["reduced"] % 256 == 0
        assert off["flags"] >= off["reduced"] + 4 * n and off["flags"] % 8 == 0
        assert off["norm_parts"] - off["flags"] >= 2 * _lib.PB_PEER_MAX * 8          # two barrier channels
        assert off["epoch"] - off["state"] >= 2 * _lib.PB_PEER_MAX * 64              # two parity slots
        assert off["total"] > off["epoch"] + 16
        for world in (1, 2, 4, 8):
            sl = lib.pb_peer_slice(n, world)
            assert sl % 4 == 0 and sl * world >= n and sl * (world - 1) < n + 4 * world


def test_argument_errors_do_not_touch_the_gpu():
    from prism_b200 import _lib
    lib = _lib.load()
    t = _lib.pb_tree()                      # null pointers
    assert lib.pb_tree_init(C.byref(t), None) == -1
    t.sum = t.min = t.state = 1
    t.capacity, t.size = 100, 100           # not a power of two
    assert lib.pb_tree_init(C.byref(t), None) == -2
    assert lib.pb_ids_select(4, 40, 10, 8, 1, 1, 0.1, 1e-10, 0.25, 1, None, None) == -3   # A > 32
    # the round's newer entry points validate before they launch, too (PB_E_ARG = -1, PB_E_UNSUPPORTED = -3)
    assert lib.pb_tc_gemm(1, 1, 128, 64, 32, None, 0, 32, 0, None, 0, 32, 0, None, 0, 0, None, 0, 0, None, 64, 0,
                          None, 0, 0, None) == -1                                       # null operands
    assert lib.pb_tc_gemm(1, 1, 128, 64, 30, 16, 0, 30, 0, 16, 0, 30, 0, None, 0, 0, None, 0, 0, 16, 64, 0,
                          None, 0, 0, None) == -3                                       # row stride not 16-byte aligned
    assert lib.pb_tc_gemm(2, 3, 128, 64, 32, 16, 0, 32, 0, 16, 0, 32, 0, None, 0, 0, None, 0, 0, 16, 64, 0,
                          None, 0, 0, None) == -1                                       # kbatches > 1 needs batch == 1
    assert lib.pb_tc_gemm_supported(128, 64, 32, 32, 32, 64) == 1 and lib.pb_tc_gemm_supported(128, 64, 32, 30, 32, 64) == 0
    assert lib.pb_layer_norm_supported(10, 3136) == 1 and lib.pb_layer_norm_supported(10, 3138) == 0
    assert lib.pb_layer_norm_supported(10, 8192) == 0
    assert lib.pb_layer_norm_fwd(8, 6, 1e-5, 16, None, None, 16, None, None, None) == -1  # F % 4 != 0
    assert lib.pb_relu_bwd_bias(1, 8, 6, 16, None, None, None, 16, None) == -1            # N % 4 != 0
    assert lib.pb_iqn_phi_bwd(4, 8, 6, 16, 16, 16, 16, None, 16, None) == -1              # F % 4 != 0
    assert lib.pb_theil_fwd(2, 65, 8, 16, 16, 16, 16, None) == -1                         # K > 64
    assert lib.pb_theil_chunks(100) == 8 and lib.pb_theil_chunks(262144) == 64 and lib.pb_theil_chunks(1 << 30) == 128
    assert lib.pb_peer_barrier(None, None) == -1
    g = _lib.pb_peer_group()
    g.world, g.rank = 9, 0                                                                # more than PB_PEER_MAX ranks
    assert lib.pb_peer_barrier(C.byref(g), None) == -1
    g.world, g.rank = 2, 0                                                                # flags / epoch missing
    assert lib.pb_peer_reduce_scatter(C.byref(g), 1024, 16, None, None) == -1
    assert lib.pb_store_scatter_dbuf(None, 4, 16, 16, 16, 16, 16, 16, 16, None) != 0
    assert lib.pb_select_copy_f64(None, 16, 16, 16, 0, 4, None) == -1
    assert lib.pb_copy_h2d_async(None, 16, 4, None) == -1


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle\.|/oracle/", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "prism_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                code = "\n".join(l for l in text.splitlines() if "oracle/" not in l or "import" in l)
                assert not re.search(r"^\s*(from|import)\s+oracle\b", code, re.M), os.path.join(dirpath, f)


def test_product_fails_loudly_without_cuda():
    import torch
    import prism_b200
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(Exception):
        prism_b200.PrioritizedTree(128, device="cpu")
    with pytest.raises(Exception):
        prism_b200.build_exp_buffer(prism_b200.minatar_dqn_per_config(device="cpu"))
    from prism_b200.agents import ops
    with pytest.raises(Exception):
        ops.cos_basis(torch.rand(8), 64)
