"""GPU diagnostic: encode every parity case on the device, compare against the oracle per stream."""
import sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import cases, refbind
from dsrc_b200 import BlockCompressor

ORDER = [0, 1, 3, 2]   # meta | tags | quality | dna
NAMES = ['meta', 'tag', 'dna', 'quality']
only = sys.argv[1:] 
bad = 0
for name, data, d, q, pr in cases.small_cases():
    if only and not any(o in name for o in only): continue
    chunk = data[:-1]
    ora = refbind.Oracle(33, pr, d, q)
    try:
        bc = BlockCompressor(33, bool(pr), d, q, max_block_bytes=max(len(chunk) + 64, 1 << 16))
    except Exception as e:
        print(name, 'CREATE FAIL', e); bad += 1; continue
    for it in range(2):
        exp, eraw, ecmp = ora.store(chunk)
        try:
            got, graw, gcmp = bc.store(chunk)
        except Exception as e:
            print(name, it, 'ENCODE FAIL', e); bad += 1; break
        if got == exp and graw == eraw and gcmp == ecmp:
            print(name, it, 'OK', len(got), bc.kernel_times() if it == 1 else '')
            continue
        bad += 1
        print(name, it, 'MISMATCH sizes', len(got), len(exp), 'comp', gcmp, ecmp, 'raw', graw == eraw)
        pe = pg = 0
        for s in ORDER:
            a = got[pg:pg + gcmp[s]]; b = exp[pe:pe + ecmp[s]]
            if a != b:
                k = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
                print('   stream', NAMES[s], 'differs at', k, 'of', len(a), len(b), a[max(0,k-4):k+8].hex(), b[max(0,k-4):k+8].hex())
            pg += gcmp[s]; pe += ecmp[s]
    bc.close()
print('BAD', bad)
