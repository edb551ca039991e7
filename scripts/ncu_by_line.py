"""Join an ncu SASS source page (csv) with nvdisasm -gi line info and aggregate samples / instructions per source line.
usage: ncu_by_line.py <sass.csv from ncu --page source --csv> <nvdisasm -gi -c output> <mangled kernel substring> [top]"""
import csv, re, sys, collections
sass_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
# nvdisasm: offset -> (innermost file:line, outermost kernel-file line)
loc = {}
cur = None; outer = None; infn = False; prev_was_loc = False
for ln in open(dis):
    if ln.startswith('.text.'):
        infn = kname in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:   # a run of //## lines = the inline chain, innermost first, kernel body last
        here = (m.group(1).split('/')[-1], int(m.group(2)))
        if not prev_was_loc: cur = here
        outer = here; prev_was_loc = True
        continue
    prev_was_loc = False
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: loc[int(m.group(1), 16)] = (cur, (outer,), m.group(2).strip())
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, isamp, iins = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
base = None
by_line = collections.Counter(); ins_line = collections.Counter(); by_outer = collections.Counter(); ins_outer = collections.Counter()
tot_s = tot_i = 0
for r in rows[2:]:
    if len(r) != len(hdr) or r[ia] == 'Address': continue
    a = int(r[ia], 16)
    if base is None: base = a
    off = a - base
    s, i = int(r[isamp] or 0), int(r[iins] or 0)
    tot_s += s; tot_i += i
    c, o, txt = loc.get(off, (None, (), ''))
    by_line[c] += s; ins_line[c] += i
    key = o[-1] if o else c        # line in the kernel body (outermost inline site)
    by_outer[key] += s; ins_outer[key] += i
print('total samples', tot_s, 'warp instructions', tot_i)
print('--- by innermost source line')
for k, v in by_line.most_common(top): print('%5.1f%% smp %5.1f%% ins  %s' % (100 * v / tot_s, 100 * ins_line[k] / tot_i, k))
print('--- by line of the kernel body (outermost inline site)')
for k, v in by_outer.most_common(30): print('%5.1f%% smp %5.1f%% ins  %s' % (100 * v / tot_s, 100 * ins_outer[k] / tot_i, k))
