"""Turn a .ncu-rep (ncu --set full) into the compact per-kernel text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/r01_attn2.ncu-rep > profiles/r01_attn_ncu.txt
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    # sm__pipe_tensor_cycles_active counts the tensor pipe for BOTH tcgen05.mma (UTCHMMA) and mma.sync (HMMA) issue
    # paths on sm_100 — it is the utilisation figure to quote for the tcgen05 kernels (the per-op
    # sm__ops_path_tensor_op_hmma_* counters stay 0 for UTCHMMA and are left out)
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (tcgen05.mma and mma.sync)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "  half-precision sub-pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall samples: long_scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall samples: short_scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "stall samples: barrier"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall samples: math_pipe_throttle"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall samples: wait"),
    ("smsp__pcsamp_warps_issue_stalled_mio_throttle", "stall samples: mio_throttle"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall samples: not_selected"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "stall samples: selected"),
    ("smsp__pcsamp_warps_issue_stalled_sleeping", "stall samples: sleeping"),
    ("smsp__pcsamp_warps_issue_stalled_tex_throttle", "stall samples: tex_throttle"),
    ("smsp__pcsamp_warps_issue_stalled_lg_throttle", "stall samples: lg_throttle"),
]


def main() -> None:
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
    rows = list(csv.reader(io.StringIO(out.stdout)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# source: {rep} (ncu --set full --clock-control none); one block per captured launch")
    for r in rows[2:]:
        print(f"\n== {r[col['Kernel Name']].strip()}")
        for key, label in WANT:
            if key in col and r[col[key]] != "":
                print(f"  {label:58s} {r[col[key]]} {units[col[key]]}")


if __name__ == "__main__":
    main()
