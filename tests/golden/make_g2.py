"""Extracts golden vector G2 from the reference notebook's committed cell output:
500 ticks of (num_active_foragers, num_active_resources, life_expectancy) printed by
examples/many_agent_systems/foraging_agents/EcoEvoJax/sim.ipynb (PRNGKey(0)).
Run in the build container (needs /root/reference): python tests/golden/make_g2.py"""
import json
import os

NB = "/root/reference/examples/many_agent_systems/foraging_agents/EcoEvoJax/sim.ipynb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "g2_foraging.json")

nb = json.load(open(NB))
ticks = []
for c in nb["cells"]:
    for o in c.get("outputs", []):
        t = "".join(o.get("text", []))
        if "num_active_foragers" not in t:
            continue
        lines = [ln.strip() for ln in t.splitlines()]
        i = 0
        while i < len(lines):
            if lines[i].isdigit() and i + 3 < len(lines) and lines[i + 1].startswith("num_active_foragers"):
                ticks.append([int(lines[i + 1].split(":")[1]), int(lines[i + 2].split(":")[1]),
                              float(lines[i + 3].split(":")[1])])
                i += 4
            else:
                i += 1
assert len(ticks) == 500, len(ticks)
json.dump({"source": "reference EcoEvoJax/sim.ipynb cell output (driver cell: PRNGKey(0), arena 2000x500, "
                     "30 neurons, 1000/500 foragers, 4000/3000 resources, 500 ticks)",
           "columns": ["num_active_foragers", "num_active_resources", "life_expectancy"], "ticks": ticks},
          open(OUT, "w"))
print("wrote", OUT, ticks[:8], ticks[-1])
