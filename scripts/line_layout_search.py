import itertools, sys
from collections import defaultdict
def wavefronts(addrs):
    banks = defaultdict(set)
    for a in addrs: banks[a % 16].add(a)
    return max(len(v) for v in banks.values())
def evaluate(N, B, RSy, RSz, CS, al, be, flips):
    N2 = N*N; NT = B*N2
    def A(c, x, y, z): return c*CS + ((x + al*y + be*z) % N) + RSy*y + RSz*z
    tot = [0,0,0]; ideal = 0
    for w in range((NT + 31)//32):
        lanes = range(32*w, min(32*w+32, NT))
        for m in range(N):
            a0 = []; a1 = []; a2 = []
            for t in lanes:
                c, ab = divmod(t, N2)
                p = [(ab % N, ab // N), (ab // N, ab % N)]
                a, b = p[flips[0]]; a0.append(A(c, m, a, b))
                a, b = p[flips[1]]; a1.append(A(c, a, m, b))
                a, b = p[flips[2]]; a2.append(A(c, a, b, m))
            tot[0] += wavefronts(a0); tot[1] += wavefronts(a1); tot[2] += wavefronts(a2)
            ideal += (len(lanes) + 15)//16
    return [t/ideal for t in tot]
N = int(sys.argv[1]); B = 8
best = []
for RSy in (N, N+1, N+2):
    for RSz in range(N*RSy, N*RSy+6):
        for pad in range(0, 3):
            CS = N*RSz + pad
            if CS > (N|1)*N*N*1.12: continue
            for al in range(N):
                for be in range(N):
                    rs = []
                    for d in range(3):
                        r0 = None
                    # evaluate flips independently per direction: compute each direction for both flips
                    r_a = evaluate(N, B, RSy, RSz, CS, al, be, (0,0,0))
                    r_b = evaluate(N, B, RSy, RSz, CS, al, be, (1,1,1))
                    r = [min(x,y) for x,y in zip(r_a, r_b)]
                    fl = [0 if x<=y else 1 for x,y in zip(r_a, r_b)]
                    best.append((sum(r), CS, RSy, RSz, al, be, fl, r))
best.sort(key=lambda t: (round(t[0],3), t[1]))
for b in best[:10]: print("sum %.3f CS %d RSy %d RSz %d alpha %d beta %d flips %s" % b[:7], ["%.2f" % v for v in b[7]])
