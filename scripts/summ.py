import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print(json.dumps({"n":d["n_gpus"],"wl":d["config"].get("baseline_config"),"value":round(d["value"],1),"ms":round(d["ms_per_step"],4),"e2e":round(d["e2e"]["value"],1),"e2e_ms":round(d["e2e"]["ms_per_step"],2),"k_ms":round(d["roofline"]["kernel_ms"],4),"frac":round(d["roofline"]["frac"],3),"splats":d["config"]["splats"],"chunks":d["config"]["chunks"],"nn":d["config"]["non_null_chunks"],"clk":d["clocks"]}))
