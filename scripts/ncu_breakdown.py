"""Per-region instruction / stall-sample breakdown of k_render_rows_ws2<4> from an ncu --set full --import-source on
capture: joins `ncu --page source --print-source sass --csv` (per-SASS-instruction counters) with the line table of
`nvdisasm -g` on the same cubin.  Usage: ncu_breakdown.py <sass.csv> <nvdisasm -g output> <tiles in the capture>"""
import collections
import csv
import re
import sys

FN = '_ZN3acb17k_render_rows_ws2ILi4EEEvNS_12RenderParamsE'
sass_csv, lines_txt, tiles = sys.argv[1], sys.argv[2], int(sys.argv[3])
src = open('ascii-chat_b200/csrc/render_dev.cuh').read().splitlines()


def fn_of_line(n):  # enclosing device function / kernel of a render_dev.cuh line
    for i in range(n - 1, -1, -1):
        if src[i].startswith(' '):
            continue
        m = re.match(r'struct\s+(\w+)', src[i])
        if m:
            return 'struct ' + m.group(1)
        m = re.search(r'\b(k_\w+|\w+)\s*\((?:const |int |uint|S |class|[A-Z])', src[i])
        if m and re.match(r'(?:__device__|__global__|template|static)', src[i]):
            return m.group(1)
    return '?'


cur, infn, amap = None, False, {}
for ln in open(lines_txt):
    if ln.startswith('.text.' + FN + ':'):
        infn = True
        continue
    if infn and ln.startswith('.text.'):
        break
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m:
        amap[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
iA, iI, iS = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = int(rows[2][iA], 16)
per, samp, tot, ts = collections.Counter(), collections.Counter(), 0, 0
for r in rows[2:]:
    k = amap.get(int(r[iA], 16) - base)
    n, s = int(r[iI]), int(r[iS])
    tot += n
    ts += s
    if k is None:
        g = 'unmapped'
    elif k[0] != 'render_dev.cuh':
        g = k[0]
    else:
        g = fn_of_line(k[1])
    per[g] += n
    samp[g] += s
print("k_render_rows_ws2<EM_HB_TRUE>: %d warp-instructions, %d tiles -> %.0f per tile; %d stall samples" % (tot, tiles, tot / tiles, ts))
print("%-28s %12s %7s %9s %12s" % ("function (render_dev.cuh)", "warp-instr", "share", "samples", "instr/tile"))
for k, v in per.most_common():
    if v:
        print("%-28s %12d %6.1f%% %8.1f%% %12.0f" % (k, v, 100 * v / tot, 100 * samp[k] / ts, v / tiles))
