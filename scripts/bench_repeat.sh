for i in 1 2 3; do python bench.py > gpurun_out/r08rep_$i.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/r08rep_$i.json")); c=d["configs"]
print("run $i: value %.3f e2e %.3f frac %.3f autoreset %.2f hostpool %.2f cfg3 %.2f cfg4 %.3f mann %.2f cfg5 %.2f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["frac"], d["with_autoreset"]["value"]/1e6, d["with_autoreset_host_pool"]["value"]/1e6, c["cfg3_share_512_envs_1gpu"]["value"]/1e6, c["cfg4_8x8_1024_envs_yaw_induction"]["value"]/1e6, c["cfg2_mann"]["value"]/1e6, c["cfg5_multi_agent_4x2_2048_envs"]["value"]/1e6))
PY
done
