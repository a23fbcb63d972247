import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"],1), round(d["e2e"]["value"],1), d["config"]["recovered_planted_shifts"], " ".join(f"{k}={v['ms']}" for k,v in d["roofline"]["kernels"].items()))
