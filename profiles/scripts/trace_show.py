import numpy as np, sys
ev = np.load('/root/repo/gpurun_out/t16_trace.npy')
def dec(r):
    e = ev[r]; e = e[e != 0]
    return (e >> 8), (e & 255)
step = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for r in (0, 1):
    tw, cw = dec(r)
    starts = np.where(cw == 1)[0]
    print('worker slot', r, 'mean step period', np.diff(tw[starts]).mean())
    s, e = starts[step], starts[step + 1]
    print([(int(c), int(t - tw[s])) for t, c in zip(tw[s:e], cw[s:e])])
tw, cw = dec(0); base = tw[np.where(cw == 1)[0][step]]
for r in (2, 3):
    ti, ci = dec(r)
    i1 = np.where(ci == 1)[0]
    per = np.diff(ti[i1])
    print('issuer', r - 2, 'median chunk period', np.median(per), 'chunks traced', len(i1))
    k0 = 16 * step - (8 if r == 3 else 0)
    for k in range(k0, k0 + 18):
        a, b = i1[k], i1[k + 1]
        print('  chunk', k, 'top', int(ti[a] - base), [(int(ci[i]), int(ti[i] - ti[a])) for i in range(a + 1, b)])
