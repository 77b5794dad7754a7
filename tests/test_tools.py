"""CPU: the profile post-processing tools run on the committed ncu launch list (keeps profiles/ reproducible)."""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_launch_summary_on_committed_capture():
    csv_path = os.path.join(ROOT, "profiles", "r01_launches_117m_v6.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), csv_path, "1"],
                         capture_output=True, text=True, check=True).stdout
    head = out.splitlines()[0]
    assert head.startswith("327 launches"), head                      # one training step of the device-resident arm
    rows = {ln.split("`")[1]: ln for ln in out.splitlines() if ln.startswith("| ") and "`" in ln}
    for k in ("attn_bwd_dkv_kernel<1, 0>", "attn_fwd_tc_kernel<1, 0>", "attn_bwd_dq_kernel<1, 0>", "gemm_tc_kernel<256, 2>",
              "loss_tv_band_kernel<__nv_bfloat16, 8>", "frontend_fwd_mma_kernel", "adamw_kernel"):
        assert k in rows, k
    share = sum(float(rows[k].split("|")[3].strip().rstrip("%")) for k in rows if k.startswith("attn_"))
    assert 70.0 < share < 85.0                                        # attention = 77.8 % of the profiled step
