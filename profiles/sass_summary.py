"""Per-kernel counts of the SASS instructions that prove the Blackwell-native paths (profiles/sass_r2.txt):
    cuobjdump -sass howtotrainyourmamlpytorch_b200/lib/libmaml_b200.so > /tmp/sass.txt && python profiles/sass_summary.py /tmp/sass.txt"""
import collections
import re
import subprocess
import sys

PAT = collections.OrderedDict([
    ("UTCHMMA", r"\bUTCHMMA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UTCBAR", r"\bUTCBAR"),
    ("UTCATOM/ALLOC", r"\bUTCATOMSWS|\bUTCALLOC"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("FFMA", r"\bFFMA"),
    ("HMMA (legacy)", r"\bHMMA"), ("LD/ST .SYS (peer)", r"\.SYS"), ("MEMBAR.SYS", r"MEMBAR\.SC\.SYS|MEMBAR\.ALL\.SYS"),
    ("UCGABAR (cluster barrier)", r"UCGABAR")])


def main():
    cur, stats = None, collections.OrderedDict()
    for line in open(sys.argv[1]):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            stats[cur] = collections.Counter()
            continue
        if cur:
            for k, p in PAT.items():
                if re.search(p, line):
                    stats[cur][k] += 1
            if re.search(r"^\s+/\*[0-9a-f]{4}\*/", line):
                stats[cur]["instructions"] += 1
    names = subprocess.run(["c++filt"] + list(stats.keys()), capture_output=True, text=True).stdout.strip().split("\n")
    tot = collections.Counter()
    for (k, c), name in zip(stats.items(), names):
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
        tot.update({x: c[x] for x in PAT})
        if any(c[x] for x in ("UTCHMMA", "LDTM", "UTMALDG", "LD/ST .SYS (peer)", "UCGABAR (cluster barrier)")):
            print("%-34s instr %5d  " % (name[:34], c["instructions"]) + "  ".join("%s %d" % (x, c[x]) for x in PAT if c[x] and x != "FFMA"))
    print("\nwhole library: " + "  ".join("%s %d" % (x, tot[x]) for x in PAT))


if __name__ == "__main__":
    main()
