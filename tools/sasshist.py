"""Scratch: SASS instruction histogram by source line for one kernel of libfreerl_b200.so (needs -lineinfo)."""
import re, collections, subprocess, sys, os, tempfile
so = sys.argv[2] if len(sys.argv) > 2 else 'freerl_b200/libfreerl_b200.so'
pat = sys.argv[1] if len(sys.argv) > 1 else 'AcAlgo'
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if f.endswith('.cubin')][0]
txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, cub)], capture_output=True, text=True).stdout
hist = collections.Counter(); fn = None; cur = None
for ln in txt.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m: fn = m.group(1); continue
    if fn is None or pat not in fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln): hist[cur] += 1
print('total instr', sum(hist.values()))
byfile = collections.Counter()
for k, v in hist.items():
    if k: byfile[k[0]] += v
print(byfile.most_common(8))
for k, v in hist.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 40):
    src = ''
    try:
        src = open('freerl_b200/csrc/' + k[0]).read().splitlines()[k[1] - 1].strip()[:110]
    except Exception: pass
    print(v, k, src)
