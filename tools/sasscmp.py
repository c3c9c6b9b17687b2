"""Per-kernel SASS comparison of two builds of libmvicp.so (instruction text, addresses and encodings ignored):
  python tools/sasscmp.py <old.so> <new.so>
lists kernels whose code changed and kernels that are new -- the proof that a change left the default device path untouched."""
import re, subprocess, sys, collections
def load(lib):
    out=subprocess.run(['cuobjdump','-sass',lib],capture_output=True,text=True).stdout
    d=collections.defaultdict(list); name=None
    for l in out.split('\n'):
        m=re.search(r'Function : (\S+)',l)
        if m: name=m.group(1); continue
        m=re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/',l)
        if m and name: d[name].append(m.group(1))
    return d
a=load(sys.argv[1]); b=load(sys.argv[2])
print(len(a),'kernels before,',len(b),'after;', sum(map(len,a.values())),'instructions before')
diff=[k for k in a if a[k]!=b.get(k)]
new=[k for k in b if k not in a]
print('changed:',diff); print('new:',[k[:60] for k in new])
