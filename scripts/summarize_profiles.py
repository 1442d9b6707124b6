"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <launches.csv> <full.ncu-rep> <tag> [<estep_traffic.csv>]
"""
import collections, csv, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1:4]
traffic = sys.argv[4] if len(sys.argv) > 4 else None

# ---- launch list: per-kernel totals and shares -----------------------------------------------------------------------
with open(launches) as f:
	lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
out_csv = os.path.join(ROOT, 'profiles', '%s_launches.csv' % tag)
with open(out_csv, 'w') as f:
	w = csv.writer(f)
	w.writerow(['id', 'kernel', 'grid', 'block', 'duration_ms'])
	for r in rows:
		w.writerow([r['ID'], r['Kernel Name'].split('(')[0].replace('void ', '').replace('trlda::', ''), r['Grid Size'], r['Block Size'],
			'%.4f' % (float(r['Metric Value'].replace(',', '')) / 1e6)])
total = collections.Counter()
count = collections.Counter()
for r in rows:
	name = r['Kernel Name'].split('(')[0].replace('void ', '').replace('trlda::', '')
	total[name] += float(r['Metric Value'].replace(',', '')) / 1e6
	count[name] += 1
grand = sum(total.values())
summary = ['# %s — ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras`' % tag,
	'(`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)', '',
	'| kernel | launches | total ms | share | mean ms |', '|---|---:|---:|---:|---:|']
for name, ms in total.most_common():
	summary.append('| `%s` | %d | %.2f | %.1f %% | %.3f |' % (name, count[name], ms, 100 * ms / grand, ms / count[name]))
summary.append('| all | %d | %.2f | 100 %% | |' % (len(rows), grand))

# ---- full capture: key metrics per captured launch --------------------------------------------------------------------
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
data = list(csv.reader(raw.splitlines()))
hdr, units, body = data[0], data[1], data[2:]
idx = {h: i for i, h in enumerate(hdr)}
keys = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
	'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
	'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
	'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
	'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
	'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
	'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic']
with open(os.path.join(ROOT, 'profiles', '%s_ncu_full.csv' % tag), 'w') as f:
	w = csv.writer(f)
	w.writerow(['metric', 'unit'] + ['launch%d' % i for i in range(len(body))])
	for k in keys:
		if k in idx:
			w.writerow([k, units[idx[k]]] + [r[idx[k]] for r in body])
summary += ['', '## `ncu --set full` capture (%s_ncu_full.csv)' % tag, '',
	'| launch | kernel | ms | DRAM read GB | DRAM write GB | DRAM %% of peak | L2 hit %% | issue active %% | regs |', '|---|---|---:|---:|---:|---:|---:|---:|---:|']
for i, r in enumerate(body):
	g = lambda k: r[idx[k]] if k in idx else ''
	summary.append('| %d | `%s` | %.3f | %.3f | %.3f | %.1f | %.1f | %.1f | %s |' % (
		i, g('Kernel Name').split('(')[0].replace('void ', '').replace('trlda::', ''), float(g('gpu__time_duration.sum')),
		float(g('dram__bytes_read.sum')), float(g('dram__bytes_write.sum')),
		float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')), float(g('lts__t_sector_hit_rate.pct')),
		float(g('smsp__issue_active.avg.pct_of_peak_sustained_active')), g('launch__registers_per_thread')))
# ---- E-step launch by launch (metrics pass over the E-step kernels only) ------------------------------------------------
if traffic:
	import json, shutil
	shutil.copyfile(traffic, os.path.join(ROOT, 'profiles', '%s_estep_traffic.csv' % tag))
	with open(traffic) as f:
		trows = list(csv.DictReader([l for l in f if not l.startswith('==')]))
	per = collections.defaultdict(dict)
	names = {}
	for r in trows:
		per[int(r['ID'])][r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
		names[int(r['ID'])] = r['Kernel Name'].split('(')[0].replace('void ', '').replace('trlda::', '')
	scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1, '%': 1}
	# one E-step call = the launches up to the next k_estep_stream launch (the documents too long for the TMEM tile come first)
	calls, current = [], None
	for i in sorted(per):
		if current is None or names[i].startswith('k_estep_stream'):
			current = []
			calls.append(current)
		current.append(i)
	summary += ['', '## E-step, call by call (%s_estep_traffic.csv)' % tag, '',
		'`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,'
		'l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:k_estep -c 70 python bench.py --steps 2 --warmup 1 --no-extras`:',
		'three update_parameters steps of ten trust-region iterations each; one E-step call = one k_estep_stream launch for the',
		'documents of more than 192 pairs + ONE persistent k_estep_tmem launch for all the others (the tile shape is picked per',
		'document).  The tile is read from HBM once per call whatever the number of inner iterations; it then lives in tensor',
		'memory.  (Multi-pass metric collection: the ms column is a replay pass with flushed caches, not a bench time.)', '',
		'| E-step call | launches | ms | HBM read GB | HBM write GB | L2 -> SM GB |', '|---:|---:|---:|---:|---:|---:|']
	tot = [0., 0., 0., 0.]
	for c, ids in enumerate(calls):
		t = rd = wr = l2 = 0.
		for i in ids:
			m = per[i]
			v = lambda k: m[k][0] * scale.get(m[k][1], 1)
			t += v('gpu__time_duration.sum'); rd += v('dram__bytes_read.sum'); wr += v('dram__bytes_write.sum'); l2 += v('l1tex__m_xbar2l1tex_read_bytes.sum')
		tot = [tot[0] + t, tot[1] + rd, tot[2] + wr, tot[3] + l2]
		summary.append('| %d | %d | %.2f | %.2f | %.2f | %.2f |' % (c, len(ids), t * 1e3, rd / 1e9, wr / 1e9, l2 / 1e9))
	n = len(calls)
	summary.append('| mean | | %.2f | %.2f | %.2f | %.2f |' % (tot[0] / n * 1e3, tot[1] / n / 1e9, tot[2] / n / 1e9, tot[3] / n / 1e9))
	with open(os.path.join(ROOT, 'profiles', '%s_estep_traffic.json' % tag), 'w') as f:
		json.dump({'cfg3/mixed': {
			'bytes_per_estep': (tot[1] + tot[2]) / n, 'l2_to_sm_bytes_per_estep': tot[3] / n, 'estep_calls': n,
			'source': 'profiles/%s_estep_traffic.csv: dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one E-step call, mean over %d calls (three steps of the bench command; cold and warm calls move the same bytes)' % (tag, n)}}, f, indent=1)

with open(os.path.join(ROOT, 'profiles', '%s_summary.md' % tag), 'w') as f:
	f.write('\n'.join(summary) + '\n')
print('\n'.join(summary))
