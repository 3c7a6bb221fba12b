"""Compare the SASS of every kernel in two builds of libsfhcuda.so (cuobjdump -sass, address comments stripped).
usage: sass_compare.py OLD.so NEW.so -- used to show that changes made without a GPU left the verified kernels byte-identical."""
import subprocess, sys, re, hashlib
def funcs(so):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    res, name, buf = {}, None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name: res[name] = hashlib.md5("\n".join(buf).encode()).hexdigest()
            name, buf = m.group(1), []
        elif name:
            buf.append(re.sub(r"/\*[0-9a-f]{4,}\*/", "", line).strip())
    if name: res[name] = hashlib.md5("\n".join(buf).encode()).hexdigest()
    return res
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
common = set(a) & set(b)
diff = [n for n in common if a[n] != b[n]]
print(f"old {len(a)} kernels, new {len(b)}; common {len(common)}; changed {len(diff)}; only-new {len(set(b)-set(a))}; only-old {len(set(a)-set(b))}")
for n in diff[:20]: print("CHANGED", n)
for n in sorted(set(a)-set(b))[:10]: print("ONLY OLD", n)
