"""debug probe: repeat the scan parity check of golden cases (intermittent-failure hunting)"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import caselib
names = sys.argv[1].split(',')
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for name in names:
    case = caselib.load_case(name)
    ctx = caselib.make_ctx(case)
    bad = 0
    for k in range(reps):
        try:
            caselib.check_scan(case, ctx)
        except AssertionError as e:
            bad += 1
            if bad == 1:
                print(name, 'FAIL', str(e)[:120].replace('\n', ' '))
    print(name, 'failures', bad, 'of', reps)
