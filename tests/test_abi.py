"""CPU: the structs of include/offt.h have the layout the reference's own offt.h gives its callers (flag set
-DA2AV -DSTRIDE, reference Makefile:21-24), so a driver compiled against either header can be linked with the
library; and the shared library exports every symbol include/*.h declares."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
PROBE = r'''
#include <stdio.h>
#include <stddef.h>
#include "offt.h"
#define P(T, f) printf(#T "." #f " %zu\n", offsetof(struct T, f))
int main(void) {
  printf("sizeof_params %zu\nsizeof_comm %zu\n", sizeof(struct _offt_params), sizeof(struct _offt_comm));
  P(_offt_params, v);
  P(_offt_comm, p1); P(_offt_comm, comm1); P(_offt_comm, M1); P(_offt_comm, F1); P(_offt_comm, m1); P(_offt_comm, b1);
  P(_offt_comm, istart); P(_offt_comm, isize); P(_offt_comm, istride); P(_offt_comm, ostart); P(_offt_comm, osize); P(_offt_comm, ostride);
  P(_offt_plan, p); P(_offt_plan, is_notest); P(_offt_plan, t_init); P(_offt_plan, t); P(_offt_plan, point_database_file);
  P(_offt_plan, params); P(_offt_plan, comm); P(_offt_plan, buffers1); P(_offt_plan, pt_transpose); P(_offt_plan, p1d_x); P(_offt_plan, p1d_xy_s_list_size);
  printf("PARAM_COUNT %d GES %d _S_ %d _T2_ %d UNPACK2 %d INIT_BUFFER %d\n", PARAM_COUNT, GES, _S_, _T2_, UNPACK2, INIT_BUFFER);
  return 0;
}
'''


def _probe(tmp_path, include_dir, extra):
    src = tmp_path / "probe.c"
    src.write_text(PROBE)
    exe = tmp_path / f"probe_{include_dir.name}"
    subprocess.run(["gcc", "-std=gnu11", "-w", *extra, f"-I{include_dir}", f"-I{ROOT / 'include' / 'compat'}", str(src), "-o", str(exe)], check=True)
    return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout


@pytest.mark.skipif(not (REF / "offt.h").exists(), reason="needs /root/reference (build container only)")
def test_struct_layout_matches_reference_header(tmp_path):
    ours = _probe(tmp_path, ROOT / "include", [])
    theirs = _probe(tmp_path, REF, ["-DA2AV", "-DSTRIDE"])
    assert ours == theirs


def test_library_exports_every_declared_symbol():
    lib = ROOT / "offt_b200" / "lib" / "libofft_b200.so"
    assert lib.exists(), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    L = ctypes.CDLL(str(lib))
    names = set()
    for h in ("offt.h", "offt_b200.h"):
        text = (ROOT / "include" / h).read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b((?:offt|offtb)_\w+|print_params|ah_tuning|params_range_setup|params_set_default|params_convert|is_infeasible_point|grid_value_floor|grid_value_ceil)\s*\(", text):
            names.add(m.group(1))
    assert len(names) > 40
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
