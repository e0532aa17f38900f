"""The critical path the GPU actually took (multi-stream run), from three measurements:

  graph.dot      MAML_B200_GRAPH_DOT=<file> python scripts/trace_timeline.py                       (dependencies)
  streams trace  python scripts/trace_timeline.py --full > <file>                                  (start time of every node)
  serial trace   MAML_B200_ONE_STREAM=1 python scripts/trace_timeline.py --full > <file>           (uncontended durations)

Trace lines carry the launch tag (#n) = kernel-node order of the captured graph, so every start time is attributed to
its node.  For node v:  ready(v) = max over parents p of start(p) + dur_serial(p);  delay(v) = start(v) - ready(v)
(> 0: a parent ran longer than alone, or v waited for SMs).  Walking back from the last node through the parent that
became ready last gives the realised critical chain; the script prints its composition and where the delays sit.

    python scripts/real_critical_path.py graph.dot trace_streams.txt trace_serial.txt
"""
import collections
import json
import os
import re
import sys


def read_trace(path):
    ents = {}
    order = []
    for line in open(path):
        m = re.match(r"\s*([\d.]+) us\s+(\S+)\s+\+(-?[\d.]+)\s+#(\d+)", line)
        if not m:
            continue
        t, nm, nxt, tag = float(m.group(1)), m.group(2), float(m.group(3)), int(m.group(4))
        if nm.endswith(":end"):
            ents[tag]["end"] = t
            continue
        ents[tag] = {"start": t, "name": nm}
        order.append(tag)
    return ents, order


def main():
    dot, streams, serial = sys.argv[1:4]
    ids = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "trace_kernel_ids.json")))
    known = sorted(ids, key=len, reverse=True)
    txt = open(dot).read()
    nodes = {}
    for m in re.finditer(r'"(graph_\d+_node_(\d+))"\[[^\]]*?label="(\d+)\n([^"]*)"\]', txt, re.S):
        body = m.group(4)
        nodes[m.group(1)] = dict(idx=int(m.group(2)), kernel="\\<\\<\\<" in body,
                                 name=next((k.replace("_kernel", "") for k in known if k in body), body.split("\n")[0][:20]),
                                 grid=(re.search(r"\\<\\<\\<\\\{([^\\]*)\\\}", body) or [None, ""])[1])
    edges = re.findall(r'"(graph_\d+_node_\d+)" -> "(graph_\d+_node_\d+)"', txt)
    knodes = sorted([k for k, n in nodes.items() if n["kernel"]], key=lambda k: nodes[k]["idx"])
    tag_of = {k: i for i, k in enumerate(knodes)}           # launch tag = kernel-node creation order
    st, _ = read_trace(streams)
    se, se_order = read_trace(serial)
    # serial durations: time to the next start in the one-stream run
    dur = {}
    for a, b in zip(se_order, se_order[1:]):
        dur[a] = se[b]["start"] - se[a]["start"]
    dur[se_order[-1]] = 5.0
    pred = collections.defaultdict(list)
    for a, b in edges:
        pred[b].append(a)

    def kernel_preds(k, seen=None):           # skip memset / empty nodes
        out = []
        for p in pred[k]:
            if nodes[p]["kernel"]:
                out.append(p)
            else:
                out.extend(kernel_preds(p))
        return out

    info = {}
    for k in knodes:
        tag = tag_of[k]
        if tag not in st:
            continue
        ps = [p for p in kernel_preds(k) if tag_of[p] in st]
        ready, via = 0.0, None
        for p in ps:
            r = st[tag_of[p]]["start"] + dur.get(tag_of[p], 0.0)
            if r > ready:
                ready, via = r, p
        info[k] = dict(start=st[tag]["start"], ready=ready, via=via, delay=st[tag]["start"] - ready if via else 0.0)
    last = max(info, key=lambda k: info[k]["start"] + dur.get(tag_of[k], 0.0))
    chain = []
    k = last
    while k is not None:
        chain.append(k)
        k = info[k]["via"]
    chain.reverse()
    total = info[last]["start"] + dur.get(tag_of[last], 0.0)
    print("realised span %.0f us; chain of %d kernels; sum of their uncontended durations %.0f us; sum of delays %.0f us" % (
        total, len(chain), sum(dur.get(tag_of[k], 0.0) for k in chain), sum(max(info[k]["delay"], 0.0) for k in chain)))
    agg = collections.OrderedDict()
    for k in chain:
        key = nodes[k]["name"] + " {" + nodes[k]["grid"] + "}"
        c, d, w = agg.get(key, (0, 0.0, 0.0))
        agg[key] = (c + 1, d + dur.get(tag_of[k], 0.0), w + max(info[k]["delay"], 0.0))
    print("%-36s %4s %10s %12s" % ("kernel {grid} on the realised chain", "n", "alone us", "+delay us"))
    for key, (c, d, w) in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
        print("%-36s %4d %10.1f %12.1f" % (key, c, d, w))
    if "--chain" in sys.argv:
        for k in chain:
            print("  #%-4d %9.1f  %-20s {%s}  alone %.1f  delay %+.1f" % (tag_of[k], info[k]["start"], nodes[k]["name"], nodes[k]["grid"],
                                                                      dur.get(tag_of[k], 0.0), info[k]["delay"]))


if __name__ == "__main__":
    main()
