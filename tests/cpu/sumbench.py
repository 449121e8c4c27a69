import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], round(d["value"],2),"blk/s pbs_ms",round(d["roofline"]["avg_launch_ms"],1),"frac",round(d["roofline"]["frac"],3),d["stage_share"],"lat",round(d["latency_s_per_block"],3),d["verified"])
