"""CPU checks of bench.py's contract: the reference arm (the reference's own modules on the host cores) prints one JSON line with the
agreed keys on the fdgan_b200 arm's metric / unit / config, only rank 0 works under torchrun, and the fdgan_b200 arm
refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env, cwd=ROOT,
                          timeout=600)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--size", "64", "--steps", "2", "--warmup", "1", "--gpus", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 2
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    # the reference's own modules (baseline/_ref or /root/reference) when present, else the oracle port -- and the line says which
    from oracle import ref_step
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_step.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["sample_batch"] == d["sample_batch"] == 4 and "4-image" in d["cpu_baseline"]["sample"]      # the batch it really runs
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench
    assert d["metric"] == bench.METRIC
    assert d["config"] == bench.workload_config(16, 2, 64)      # the fdgan_b200 arm's config, key for key


def test_reference_arm_other_ranks_do_no_work():
    r = _run(["--impl", "reference", "--gpus", "2"], {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_committed_traffic_file_belongs_to_the_committed_kernels():
    """profiles/r02_traffic.json feeds roofline.traffic; bench.py refuses it (traffic = null) when the kernel sources changed after the ncu
    launch list was captured.  This test makes a stale file visible on the CPU side before a round closes."""
    import json
    import bench
    with open(os.path.join(ROOT, bench.TRAFFIC_FILE)) as f:
        js = json.load(f)
    assert js["src_sha"] == bench.kernel_source_sha(), "re-capture the launch list (tools/r2_final_profiles.sh) after the last kernel change"
    fam, why = bench.ncu_family_profile("conv_tcgen05")
    assert fam is not None and fam["dram_bytes_per_launch"] > 0, why
