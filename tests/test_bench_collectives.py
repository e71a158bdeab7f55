"""bench.py runs one process per GPU; a collective (barrier, all-reduce, the border exchange) that only SOME ranks reach
deadlocks the job until the NCCL watchdog fires.  No GPU is needed to rule that out: walk bench.py's syntax tree and
require every collective call to sit outside any branch whose condition can differ between ranks."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# calls that every rank has to make the same number of times, in the same order
COLLECTIVE_ATTRS = {"barrier", "reduce", "all_reduce", "exchange_halos", "rebuild_step", "exchange_begin", "exchange_finish", "destroy_process_group"}
COLLECTIVE_NAMES = {"pcie_floor_ms", "run_world"}
# conditions that are the same on every rank of a run (arguments, the world size, the presence of a process group)
RANK_INVARIANT = {"with_mesh_all", "border is not None", "self.dist", "comm.dist", "dist", "not args.no_extra and args.workload == 'c2'", "n == 1",
                  "args.impl == 'reference'", "world_size > 1"}


def collective_calls(tree):
    parents = {}
    for node in ast.walk(tree):
        for child in ast.iter_child_nodes(node):
            parents[child] = node
    for node in ast.walk(tree):
        if not isinstance(node, ast.Call):
            continue
        f = node.func
        name = f.attr if isinstance(f, ast.Attribute) else (f.id if isinstance(f, ast.Name) else None)
        if (isinstance(f, ast.Attribute) and name in COLLECTIVE_ATTRS) or (isinstance(f, ast.Name) and name in COLLECTIVE_NAMES):
            chain, p = [], node
            while p in parents:
                p = parents[p]
                if isinstance(p, (ast.If, ast.While, ast.IfExp)):
                    chain.append(ast.unparse(p.test))
            yield node.lineno, name, chain


def test_no_collective_inside_a_rank_dependent_branch():
    src = open(os.path.join(ROOT, "bench.py")).read()
    calls = list(collective_calls(ast.parse(src)))
    assert len(calls) >= 15                                   # the walk found bench.py's barriers and reductions
    bad = [(line, name, cond) for line, name, chain in calls for cond in chain if cond not in RANK_INVARIANT]
    assert not bad, "collective under a condition that is not known to be the same on every rank: %r" % bad


def test_slab_exchange_pairs_every_send_with_a_receive():
    """halo_schedule is the only place that decides who talks to whom: for every world size the sends of one rank are the
    receives of its neighbour, plane for plane."""
    from voxplat_b200 import slab
    for ws in (1, 2, 4, 8):
        for mesh in (False, True):
            plans = [slab.halo_schedule(r, ws, mesh) for r in range(ws)]
            for r, plan in enumerate(plans):
                for op, which, peer in plan:
                    other = "recv" if op == "send" else "send"
                    assert (other, which, r) in plans[peer], (ws, mesh, r, op, which, peer)
