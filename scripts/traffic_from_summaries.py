"""profiles/r02_ncu_traffic.json from the raw ncu summary tables (scripts/ncu_summarize.py output): dram bytes read + written
of the first launch of gram_tc2 (step-1 Gram), gram_tc (step-0 Gram) and the first two project_tc launches.

    python scripts/traffic_from_summaries.py gpurun_out/r02_ncu_summaries_raw.md profiles/r02_ncu_traffic.json"""
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
sections, cur = {}, None
for line in open(sys.argv[1]):
    if line.startswith("## "):
        cur = line[3:].strip()
        sections[cur] = []
    elif cur and line.startswith("| dram__bytes_"):
        m = re.match(r"\| (dram__bytes_\w+\.sum) \| ([0-9.,]+) (\w+) \|", line)
        sections[cur].append((m.group(1), float(m.group(2).replace(",", "")) * UNIT[m.group(3)]))


def launch(sec, i):  # i-th launch of the section: (read, write) pairs in order
    v = sections[sec]
    return v[2 * i][1] + v[2 * i + 1][1]


out = {
    "gram1": launch("gram_tc2_kernel", 0),
    "gram0": launch("gram_tc_kernel", 0),
    "factor0": launch("project_tc_kernel", 0),
    "factor1": launch("project_tc_kernel", 1),
    "source": "ncu --set full --clock-control none, 64^5 fp32 r=32, one call (scripts/gpu_r2_n.sh); "
              "dram__bytes_read.sum + dram__bytes_write.sum per launch, extracted by scripts/traffic_from_summaries.py",
}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
