"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm
(`--impl reference`: the reference's own sources behind the shims, on the host cores) prints exactly
one JSON line with the keys of the GPU arm, under torchrun only rank 0 prints, and the GPU arm refuses
to run - loudly - where there is no CUDA device (no CPU fallback for the product path)."""
import json
import os
import subprocess
import sys

import pytest

import oraclelib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def run_bench(*args, launcher=()):
    p = subprocess.run([sys.executable, *launcher, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    return p.returncode, [l for l in p.stdout.splitlines() if l.strip()], p.stderr


@needs_ref
def test_reference_arm_prints_one_line_with_the_contract_keys():
    rc, lines, err = run_bench("--impl", "reference", "--workload", "ml1m", "--steps", "1", "--warmup", "1",
                               "--cpu-sample", "200000")
    assert rc == 0, err[-2000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["metric"] == "rating updates/sec per epoch" and d["unit"] == "updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
    # a sampled run says so in its own config (it does not borrow the label of the arm it stands beside)
    assert d["config"]["workload"].startswith("first ") and "ml1m-shaped" in d["config"]["workload"] and d["config"]["k"] == 32
    assert 200000 <= d["config"]["ratings"] == d["train_ratings"] < 300000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "ratings" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@needs_ref
def test_reference_arm_under_torchrun_only_rank_zero_prints():
    rc, lines, err = run_bench("--impl", "reference", "--gpus", "2", "--workload", "ml1m", "--steps", "1", "--warmup", "1",
                               "--cpu-sample", "100000",
                               launcher=("-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                         "--master-addr", "127.0.0.1", "--master-port", "29571"))
    assert rc == 0, err[-2000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["config"]["parallelism"] == "dsgd2"


@needs_ref
def test_reference_arm_runs_the_whole_file_by_default_and_loads_no_product_library():
    """no --cpu-sample: the whole file of the workload, labelled as such; the arm's own process never maps
    libmf_b200.so (the data come from the getdata binary, a separate process)"""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'ml1m', '--steps', '1', '--warmup', '1'];"
            "runpy.run_path(%r, run_name='__main__');"
            "import os; maps = open('/proc/self/maps').read(); assert 'libmf_b200' not in maps, 'product library mapped'"
            % os.path.join(ROOT, "bench.py"))
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.strip()][0])
    assert d["config"]["workload"].startswith("ml1m-shaped") and d["config"]["ratings"] == 1_000_000
    assert 850_000 < d["train_ratings"] < 950_000 and "all " in d["cpu_baseline"]["sample"]


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    rc, lines, err = run_bench("--steps", "1", "--no-cpu", "--workload", "ml1m")
    assert rc != 0 and not lines
    assert "no CPU fallback" in err
