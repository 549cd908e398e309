import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print(sys.argv[1], d['value'], d['e2e']['value'], json.dumps(d['around_the_solve'].get('marginalize'))[:260])
