#!/usr/bin/env python
"""Print the per-entry-point and per-shape time table of a bench.py --profile-out JSON."""
import collections, json, sys
d = json.load(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
print('step %.2f ms, instrumented %.2f ms' % (d['step_ms'], d['instrumented_total_ms']))
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms'])[:16]:
    print('  %-28s calls %4d  %8.2f ms  %5.1f%%  %s' % (k, v['calls'], v['ms'], 100 * v['share'], ('%.0f TF' % v['tflops']) if 'tflops' in v else ''))
c = d.get('calls', {})
for name in c:
    agg = collections.OrderedDict()
    for r in c[name]:
        a = r['args']
        if name == 'gn_gemm_bf16':
            key = (a[4], a[5], a[6], 'xf' if a[14] else '', 'bn' if a[16] else '', 'sc' if a[11] else ''); fl = 2 * a[4] * a[5] * a[6]
            by = 2 * a[4] * (a[6] + a[5] * (3 if a[16] else 1))
        elif name == 'gn_gemm_tn_bf16':
            key = (a[4], a[5], a[6]); fl = 2 * a[4] * a[5] * a[6]; by = 2 * a[6] * (a[4] + a[5])
        elif name == 'gn_conv1x1_bwd_bf16':
            key = (a[4], a[5], 128, 'dgrad+bn+wgrad'); fl = 4 * a[4] * a[5] * 128; by = 2 * a[4] * (128 + a[5] * (3 if a[16] else 2))
        elif name == 'gn_conv3x3_bf16':
            key = (a[2], a[3], a[5], a[8], 'bn' if a[11] else ''); fl = 2 * 9 * a[2] * a[3] * a[4] * a[5] * a[8]
            by = 2 * a[2] * a[3] * a[4] * (a[5] + a[8] * (2 if a[11] else 1))
        else:
            key = (a[4], a[5], a[7], a[8]); fl = 2 * 9 * a[4] * a[5] * a[6] * a[7] * a[8]; by = 2 * a[4] * a[5] * a[6] * (a[7] + a[8])
        e = agg.setdefault(key, [0, 0.0, 0.0, 0.0]); e[0] += 1; e[1] += r['ms']; e[2] += fl; e[3] += by
    print(name)
    for k, (n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('   %-44s x%-3d %7.2f ms  %5.0f TF  %5.0f GB/s (algorithmic)' % (k, n, ms, fl / ms / 1e9, by / ms / 1e6))
