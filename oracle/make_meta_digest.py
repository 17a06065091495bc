"""Golden digest of the graphs the reference ships (test_v1/model/*.meta, TF 1.10.1) for tests/test_host.py: per
checkpoint, every node the `.meta` writer (dl_ofdm_b200/tfmeta.py) must reproduce -- the nodes reachable from the named
fetches of dev/py/model.py:51-72 plus the model variables with their initializers / Assign / read nodes -- hashed over
(name, op, inputs, device, every attribute).  Run here (needs /root/reference):  python oracle/make_meta_digest.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get('DL_OFDM_REF', '/root/reference')


def node_key(n):
    attrs = ''.join('%s=%s;' % (k, n.attr[k].SerializeToString(deterministic=True).hex()) for k in sorted(n.attr.keys()))
    return '%s|%s|%s|%s|%s' % (n.name, n.op, ','.join(n.input), n.device, attrs)


def needed_nodes(nodes, fetches):
    """Reachable from the fetches, + every non-optimizer variable's Assign (which pulls in its initializer)."""
    targets = list(fetches) + [n for n in nodes if n.endswith('/Assign') and 'Adam' not in n and '_power' not in n
                               and not n.startswith('save/')]
    seen, st = set(), targets
    while st:
        n = st.pop().lstrip('^').split(':')[0]
        if n not in seen:
            seen.add(n)
            st.extend(nodes[n].input)
    return sorted(seen)


def digest(graph_nodes, names):
    h = hashlib.sha256()
    for n in names:
        h.update(node_key(graph_nodes[n]).encode())
        h.update(b'\n')
    return h.hexdigest()


def main():
    from tensorboard.compat.proto import meta_graph_pb2
    from dl_ofdm_b200.tfmeta import FETCHES
    out = {}
    for nb in (1, 2, 3, 4):
        for cp in (True, False):
            name = 'OFDM_Dense3_%dmod_snr%d_cp%s' % (nb, 3 * nb, cp)
            m = meta_graph_pb2.MetaGraphDef()
            m.ParseFromString(open(os.path.join(REF, 'test_v1', 'model', name + '.meta'), 'rb').read())
            nodes = {n.name: n for n in m.graph_def.node}
            names = needed_nodes(nodes, FETCHES)
            out[name] = {'nodes': len(names), 'sha256': digest(nodes, names), 'names_sha256':
                         hashlib.sha256('\n'.join(names).encode()).hexdigest(), 'tf_version': m.meta_info_def.tensorflow_version}
            if nb == 2 and cp:
                out[name]['per_node'] = {n: hashlib.sha1(node_key(nodes[n]).encode()).hexdigest()[:12] for n in names}
            print(name, out[name]['nodes'], out[name]['sha256'][:16])
    json.dump(out, open(os.path.join(ROOT, 'tests', 'golden', 'v1_meta_digest.json'), 'w'), indent=0, sort_keys=True)


if __name__ == '__main__':
    main()
