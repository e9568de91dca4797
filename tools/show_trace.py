import json, sys
res = json.load(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/trace_attn.json'))
sel = set(int(x) for x in sys.argv[2].split(',')) if len(sys.argv) > 2 else None
for r in res:
    print(r['name'], 'setup_done', r['setup_done'], 'loop_done', r['softmax_loop_done'], 'o_read', r['o_read'], 'stored', r['stored'], 'end', r['end'])
    for t in r['tiles']:
        if t['sm0'][0] is None or (sel and t['j'] not in sel): continue
        d = lambda a: [a[0]] + [a[i] - a[i - 1] for i in range(1, len(a)) if a[i] is not None]
        print(t['j'], 'sm0', d(t['sm0']), '| sm1', d(t['sm1']))
        print('   mma0', t['mma0'], '| mma1', t['mma1'], '| K', t['ldK'], 'V', t['ldV'])
