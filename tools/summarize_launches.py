"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of ONE training
step (the launches between the last two fused_sgd kernels) -> markdown.  usage: summarize_launches.py in.csv out.md"""
import collections
import csv
import sys


def main(src, dst, title):
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    data = [row for row in r if len(row) == len(hdr)]
    sgd = [i for i, row in enumerate(data) if 'fused_sgd' in row[ix['Kernel Name']]]
    step = data[sgd[-2] + 1:sgd[-1] + 1]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in step:
        name = row[ix['Kernel Name']].split('(')[0].replace('void ', '')
        tot[name] += float(row[ix['Metric Value']]) / 1e6
        cnt[name] += 1
    total = sum(tot.values())
    with open(dst, 'w') as f:
        f.write('# %s\n\n' % title)
        f.write('One training step (ResNet-50, batch 256, 224x224, 1x B200) = the launches between the last two '
                '`fused_sgd` kernels of the ncu launch list (`gpu__time_duration.sum`, `--clock-control none`; '
                'per-launch times are cold-cache and serialised: compare SHARES).\n\n')
        f.write('| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n')
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write('| `%s` | %d | %.3f | %.1f%% |\n' % (k, cnt[k], v, 100 * v / total))
        f.write('| **sum** | %d | %.3f | 100%% |\n' % (len(step), total))
        ours = sum(v for k, v in tot.items() if k.startswith('b200::'))
        f.write('\nKernels of this repo (`b200::*`): %.1f%% of the step; the rest is torch glue (loss, fills).\n'
                % (100 * ours / total))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else 'ncu launch list summary')
