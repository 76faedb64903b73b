"""CPU: the C-ABI library loads, exports every symbol include/v3d.h declares, and the ctypes mirrors of
its structs have the layout the C compiler gives them.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from video_3d_llm_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "v3d.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(v3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = cabi.load()
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/v3d.h but not exported by libv3d.so"
    assert sorted(cabi.SYMBOLS) == names, "cabi.SYMBOLS and include/v3d.h disagree"


def test_version_and_pure_host_helpers():
    lib = cabi.load()
    assert lib.v3d_version() == 102
    assert [lib.v3d_pool_points(m) for m in range(5)] == [1, 2, 9, 5, 1]
    assert lib.v3d_pool_points(99) == -1
    assert lib.v3d_mc_hash_slots(0) == 64 and lib.v3d_mc_hash_slots(1000) == 4096
    assert lib.v3d_pe_table_bytes(3584, 1, 301, cabi.BF16) == (3 * 302 + 1) * 1208 * 2
    assert lib.v3d_pe_table_bytes(12, 1, 301, cabi.BF16) == 0  # nf < 8 is unsupported, reported as 0 bytes


def test_argument_errors_are_reported_without_a_gpu():
    lib = cabi.load()
    assert lib.v3d_discretize(None, 0, None, None, 0, 0.1, 0, cabi.F32, None, None) == cabi.V3D_OK  # empty input: nothing to do
    rc = lib.v3d_discretize(None, 4, None, None, 0, 0.1, 0, cabi.F32, None, None)
    assert rc == cabi.V3D_E_ARG and b"null pointer" in lib.v3d_last_error()
    with pytest.raises(ValueError):
        cabi.check(rc, "v3d_discretize")
    with pytest.raises(NotImplementedError):
        cabi.check(cabi.V3D_E_UNSUPPORTED)


C_PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include "v3d.h"
#define F(T, m) printf(#T "." #m " %zu\n", offsetof(T, m))
int main(void) {
  printf("sizeof.v3d_frontend_desc %zu\n", sizeof(v3d_frontend_desc));
  F(v3d_frontend_desc, depth); F(v3d_frontend_desc, depth_dtype); F(v3d_frontend_desc, poses); F(v3d_frontend_desc, intr);
  F(v3d_frontend_desc, n_frames); F(v3d_frontend_desc, H); F(v3d_frontend_desc, W); F(v3d_frontend_desc, depth_scale);
  F(v3d_frontend_desc, row_map); F(v3d_frontend_desc, col_map); F(v3d_frontend_desc, src_col_min); F(v3d_frontend_desc, src_col_max);
  F(v3d_frontend_desc, patch); F(v3d_frontend_desc, grid); F(v3d_frontend_desc, pool_mode); F(v3d_frontend_desc, has_clamp);
  F(v3d_frontend_desc, clamp_lo); F(v3d_frontend_desc, clamp_hi); F(v3d_frontend_desc, discrete); F(v3d_frontend_desc, min_xyz);
  F(v3d_frontend_desc, max_xyz); F(v3d_frontend_desc, voxel_size); F(v3d_frontend_desc, div_mode); F(v3d_frontend_desc, pe_mode);
  F(v3d_frontend_desc, E); F(v3d_frontend_desc, dim_t); F(v3d_frontend_desc, table); F(v3d_frontend_desc, table_rows);
  F(v3d_frontend_desc, dtype); F(v3d_frontend_desc, tokens_in); F(v3d_frontend_desc, pe_out); F(v3d_frontend_desc, pooled_out);
  F(v3d_frontend_desc, pe_coords_out); F(v3d_frontend_desc, depth_resampled); F(v3d_frontend_desc, coord_dtype); F(v3d_frontend_desc, exact_pooling);
  printf("sizeof.v3d_mc_scene %zu\n", sizeof(v3d_mc_scene));
  F(v3d_mc_scene, depth); F(v3d_mc_scene, poses); F(v3d_mc_scene, intr); F(v3d_mc_scene, pc_voxels); F(v3d_mc_scene, n_frames);
  F(v3d_mc_scene, n_pc); F(v3d_mc_scene, seed); F(v3d_mc_scene, hash_table); F(v3d_mc_scene, hash_slots); F(v3d_mc_scene, bitsets);
  F(v3d_mc_scene, record);
  printf("const.V3D_MC_RECORD_INTS %d\n", V3D_MC_RECORD_INTS);
  return 0;
}
"""


def test_ctypes_struct_layout_matches_the_header():
    """The header is plain C: compile a probe with gcc and compare offsets with the ctypes mirrors."""
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "probe.c"), os.path.join(d, "probe")
        with open(src, "w") as f:
            f.write(C_PROBE)
        subprocess.run(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = {}
    for line in out.splitlines():
        k, v = line.rsplit(" ", 1)
        got[k] = int(v)
    assert got["sizeof.v3d_frontend_desc"] == C.sizeof(cabi.FrontendDesc)
    assert got["sizeof.v3d_mc_scene"] == C.sizeof(cabi.McScene)
    assert got["const.V3D_MC_RECORD_INTS"] == cabi.MC_RECORD_INTS
    for name, _ in cabi.FrontendDesc._fields_:
        assert got[f"v3d_frontend_desc.{name}"] == getattr(cabi.FrontendDesc, name).offset, name
    for name, _ in cabi.McScene._fields_:
        if name != "pad_":
            assert got[f"v3d_mc_scene.{name}"] == getattr(cabi.McScene, name).offset, name
