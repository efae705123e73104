"""profiles/r2_traffic.json from an ncu launch list (gpu__time_duration + dram__bytes_read/write per launch):
DRAM bytes per launch of each library entry point, averaged over the captured launches of its kernel family.

    python tools/traffic_from_ncu.py gpurun_out/r2a_launches.csv profiles/r2_traffic.json
"""
import json
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from ncu_summary import load

FAMILY = {  # entry point -> substring of the kernel name
    "ehgr_pw_gemm_bn": "pw_gemm_tc_kernel", "ehgr_pw_gemm_w16": "pw_gemm_tc_kernel", "ehgr_pw_gemm": "pw_gemm_tc_kernel",
    "ehgr_pw_wgrad": "pw_wgrad_tc_kernel", "ehgr_dw_bwd": "dw_bwd_sw_kernel", "ehgr_dw_fwd_bn": "dw_fwd_sw_kernel",
    "ehgr_dw_fwd": "dw_fwd_sw_kernel", "ehgr_bn_bwd_reduce_fin": "bn_bwd_reduce_kernel", "ehgr_bn_bwd_reduce": "bn_bwd_reduce_kernel",
    "ehgr_row_apply": "row_apply_kernel", "ehgr_stem_fwd_bn": "stem_fwd32", "ehgr_stem_wgrad": "stem_wgrad32",
    "ehgr_action_xs": "action_xs", "ehgr_action_bwd_dxs": "action_bwd_dxs",
}


def main():
    rows = load(sys.argv[1])
    out = {}
    for entry, sub in FAMILY.items():
        sel = [r for r in rows if sub in r["name"]]
        if sel:
            out[entry] = int(sum(r.get("rd", 0) + r.get("wr", 0) for r in sel) / len(sel))
    doc = {"source": f"{sys.argv[1]} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                     "--clock-control none over eager steps of `bench.py --no-graph`; mean over the captured launches "
                     "of each kernel family)",
           "bytes_per_launch": out}
    with open(sys.argv[2], "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
