"""CPU: the parts of bench.py's contract that need no GPU - the reference arm prints exactly one JSON line with the
keys the driver reads, on a small grid; and the product never reaches for the oracle."""
import json
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "ref_dump").exists(), reason="oracle/_ref/ref_dump not built")
def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1",
                          "--grid", "64"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under offt_b200/ or include/ may import, link or execute it"""
    offenders = []
    for path in list((ROOT / "offt_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if path.is_file() and path.suffix in {".py", ".cu", ".cuh", ".h", ".c", ".cpp"} or path.name in {"Makefile", "offtrun"}:
            text = path.read_text(errors="ignore")
            if re.search(r"\boracle\b", text, flags=re.I) and "oracle" in text.lower():
                for n, line in enumerate(text.splitlines(), 1):
                    code = line.strip()
                    if code.startswith(("//", "#!", "# ", "*", "/*")):
                        continue   # prose
                    if re.search(r"(^\s*(import|from)\s+\S*oracle|#\s*include.*oracle|dlopen\(.*oracle|CDLL\(.*oracle|-loracle|oracle/)", line, flags=re.I):
                        offenders.append(f"{path.relative_to(ROOT)}:{n}: {line.strip()}")
    assert not offenders, offenders


def test_own_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert res.returncode != 0
    assert "no CUDA device" in (res.stderr + res.stdout)
    assert not res.stdout.strip().startswith("{")


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "ref_dump").exists(), reason="oracle/_ref/ref_dump not built")
def test_cpu_baseline_object():
    sys.path.insert(0, str(ROOT))
    import bench
    b = bench.cpu_baseline_of((64, 64, 64))
    assert b["kind"] == "reference" and b["unit"] == "GFLOP/s" and b["value"] > 0 and b["cores"] >= 1 and "sample" in b


def test_cpu_baseline_b_object():
    sys.path.insert(0, str(ROOT))
    import bench
    b = bench.scipy_baseline((32, 32, 32))
    assert b["unit"] == "GFLOP/s" and b["value"] > 0 and b["cores"] >= 1 and "scipy" in b["kind"]


def test_reference_arm_names_the_grid_it_ran():
    """--ref-sample forces the bounded sample: config.workload must then name the sample, not the own arm's grid"""
    if not (ROOT / "oracle" / "_ref" / "ref_dump").exists():
        pytest.skip("oracle/_ref/ref_dump not built")
    sys.path.insert(0, str(ROOT))
    import bench
    assert bench.config_of(type("A", (), {"gpus": 2})(), (512, 512, 512))["workload"].startswith("512x512x512")


def test_layout_views_match_the_oracle_helpers():
    """offt_b200.layout (what bench.py's parity gate uses) against the oracle's index arithmetic on uneven and even boxes"""
    import numpy as np
    import offt_b200 as ob
    from offt_b200 import layout
    from oracle import oracle as O
    for N, p, p1, S in [((12, 10, 9), 6, 3, 0), ((16, 8, 32), 4, 2, 1), ((8, 8, 8), 1, 1, 0)]:
        grid = O.grid_values(3, *N)
        boxes = [ob.comm_box(*N, p, p1, r, S, 0) for r in range(p)]
        alloc = ob.alloc_elems(*N, p, p1)
        arrays = []
        for r, b in enumerate(boxes):
            rb = O.RankBox(p=p, rank=r, N=N, p1=p1, p2=p // p1, istart=b["istart"], isize=b["isize"], istride=b["istride"],
                           ostart=b["ostart"], osize=b["osize"], ostride=b["ostride"], alloc=alloc)
            a = layout.scatter_input(b, grid, alloc)
            assert np.array_equal(a, O.scatter_input(rb, grid))
            arrays.append(a)
        assert np.array_equal(layout.gather_input(boxes, arrays, N), grid)
        # a spectrum scattered through the output boxes comes back through gather_output
        spec = O.grid_values(4, *N)
        outs = []
        for b in boxes:
            a = np.zeros(alloc, dtype=np.complex128)
            if min(b["osize"]) > 0:
                layout.output_view(b, a)[...] = spec[tuple(slice(s, s + n) for s, n in zip(b["ostart"], b["osize"]))]
            outs.append(a)
        assert np.array_equal(layout.gather_output(boxes, outs, N), spec)
