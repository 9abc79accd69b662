"""SASS instructions per source line of one kernel (code-size view: what fills the instruction cache).
    python profiles/sass_lines.py silent-speech-decoder_b200/build/beam.o beam_kernelILb0 [top]
Needs an object built with -lineinfo; uses cuobjdump -xelf + nvdisasm -g in a temporary directory."""
import collections, os, re, subprocess, sys, tempfile

obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", cubin], cwd=d, capture_output=True, text=True).stdout.splitlines()
heads = [i for i, l in enumerate(sass) if l.startswith(".text.")]
start = [i for i in heads if pat in sass[i]][0]
end = min([i for i in heads if i > start] + [len(sass)])
cur, cnt = None, collections.Counter()
for line in sass[start:end]:
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1), int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        cnt[cur] += 1
total = sum(cnt.values())
print(f"{sass[start]}  {total} instructions = {total * 16 / 1024:.0f} KB")
src = {}
for (f, l), c in cnt.most_common(top):
    if f not in src:
        cand = [os.path.join(r, f) for r, _, fs in os.walk(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))) if f in fs]
        src[f] = open(cand[0]).read().splitlines() if cand else []
    text = src[f][l - 1].strip()[:100] if l <= len(src[f]) else ""
    print(f"{c:6d} {100 * c / total:5.1f}%  {f}:{l}  {text}")
