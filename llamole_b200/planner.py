"""Batched A* (Retro*) expansion -- the planner side of SURVEY.md section 8f-1.

The reference's search loop (planner/molstar.py:11-76) expands ONE open molecule per iteration: one `expand_fn(mol)` call =
one LLM analysis + one B=1 GIN encoder call + one B=1 predictor call (modeling_llamole.py:784-889), and every new reactant is
scored by a separate `value_fn(mol, parent)` call inside `MolTree._add_mol_node` (mol_tree.py:27-30).  That call pattern can
never reach the batched candidate-scoring kernel (BASELINE.json configs[3]).  `molstar_batched` keeps the reference's tree
(any object with `MolTree`'s interface: `mol_nodes`, `succ`, `search_status`, `root.succ_value`, `expand`, `get_best_route`)
and its selection rule, but

  * selects the `beam` open molecules with the smallest `v_target()` per iteration and hands them to ONE
    `expand_batch_fn(list_of_smiles) -> list_of_results` call (e.g. `GraphPredictor.sample_templates_batch`: one predictor
    launch for all of them), and
  * scores all new reactants of the round with ONE `value_batch_fn(list_of_smiles) -> list_of_values` call, served to the
    tree's per-node `value_fn` from a cache.

With `beam=1` the sequence of expansions, the tree and the returned route are exactly the reference's (checked against the
verbatim `molstar` in tests/test_planner_batched.py wherever /root/reference exists).  The tree itself, the LLM calls and the
RDKit / rdchiral chemistry stay the reference's: this module holds no chemistry and no tree arithmetic.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


class ValueCache:
    """`value_fn(mol, parent)` for MolTree backed by batched evaluations: `prefetch(mols)` evaluates the molecules it has not
    seen with ONE `value_batch_fn` call; a lookup miss falls back to a batch of one."""

    def __init__(self, value_batch_fn: Callable[[List[str]], Sequence[float]]):
        self._fn = value_batch_fn
        self._cache: Dict[str, float] = {}
        self.batch_calls = 0
        self.evaluated = 0

    def prefetch(self, mols: Sequence[str]) -> None:
        todo = [m for m in dict.fromkeys(mols) if m not in self._cache]
        if not todo:
            return
        vals = list(self._fn(todo))
        if len(vals) != len(todo):
            raise ValueError(f"value_batch_fn returned {len(vals)} values for {len(todo)} molecules")
        self.batch_calls += 1
        self.evaluated += len(todo)
        for m, v in zip(todo, vals):
            self._cache[m] = float(v)

    def __call__(self, mol: str, parent=None) -> float:
        if mol not in self._cache:
            self.prefetch([mol])
        return self._cache[mol]


def _default_tree_factory():
    for mod in ("src.model.planner.mol_tree", "planner.mol_tree"):
        try:
            return __import__(mod, fromlist=["MolTree"]).MolTree
        except Exception:
            continue
    raise ImportError("molstar_batched needs the reference's planner.mol_tree.MolTree (or pass tree_factory=)")


def molstar_batched(target_mol, target_mol_id, starting_mols, expand_batch_fn: Callable[[List[str]], List[Optional[dict]]],
                    value_batch_fn: Callable[[List[str]], Sequence[float]], iterations: int, beam: int = 8, viz: bool = False,
                    viz_dir=None, max_time: float = 300, tree_factory=None, stats: Optional[dict] = None):
    """Drop-in for `molstar` (same return value: `(succ, best_route, iterations_used)`) that expands up to `beam` open
    molecules per iteration.  `expand_batch_fn(mols)` returns, per molecule, the dict `expand_fn` returns in the reference
    (`reactants`, `scores`, `templates`, `analysis`) or None; `value_batch_fn(mols)` returns one value per molecule.
    `iterations` bounds the number of EXPANDED molecules (like the reference's iteration count), `stats` (optional dict)
    receives the number of batched calls."""
    if beam < 1:
        raise ValueError("beam must be >= 1")
    MolTree = tree_factory if tree_factory is not None else _default_tree_factory()
    values = ValueCache(value_batch_fn)
    values.prefetch([target_mol])
    mol_tree = MolTree(target_mol=target_mol, known_mols=starting_mols, value_fn=values)
    expanded = 0
    expand_calls = 0
    started_without_expansion = 0   # the reference counts an iteration that stops before expanding (time-out, nothing open)
    start_time = time.time()
    done = bool(mol_tree.succ)
    while not done and expanded < iterations:
        if time.time() - start_time > max_time:
            started_without_expansion = 1
            break
        scores = np.array([m.v_target() if m.open else np.inf for m in mol_tree.mol_nodes])
        if np.min(scores) == np.inf:
            started_without_expansion = 1
            break
        mol_tree.search_status = np.min(scores)
        # the `beam` best open molecules, best first (stable: ties keep the tree's node order, like np.argmin)
        order = np.argsort(scores, kind="stable")[: min(beam, iterations - expanded)]
        batch = [mol_tree.mol_nodes[int(i)] for i in order if scores[int(i)] < np.inf]
        results = expand_batch_fn([m.mol for m in batch])
        if len(results) != len(batch):
            raise ValueError(f"expand_batch_fn returned {len(results)} results for {len(batch)} molecules")
        expand_calls += 1
        # one batched evaluation for every reactant that the round's expansions will add to the tree
        new_mols: List[str] = []
        for r in results:
            if r is not None and len(r["scores"]) > 0:
                for rs in r["reactants"]:
                    new_mols.extend(set(rs.split(".")))
        values.prefetch(new_mols)
        for m_next, result in zip(batch, results):
            if not m_next.open:      # closed by an earlier expansion of this round (it cannot happen in the reference's tree; kept as a guard)
                continue
            expanded += 1
            if result is not None and len(result["scores"]) > 0:
                costs = 0.0 - np.log(np.clip(np.array(result["scores"]), 1e-3, 1.0))
                reactant_lists = [list(set(rs.split("."))) for rs in result["reactants"]]
                succ = mol_tree.expand(m_next, reactant_lists, costs, result["templates"], result["analysis"])
                if succ or mol_tree.root.succ_value <= mol_tree.search_status:    # solved / found the optimal route
                    done = True
                    break
            else:
                mol_tree.expand(m_next, None, None, None, None)
    best_route = None
    if mol_tree.succ:
        best_route = mol_tree.get_best_route()
        assert best_route is not None
    if stats is not None:
        stats.update(expand_calls=expand_calls, expanded=expanded, value_batch_calls=values.batch_calls, values_evaluated=values.evaluated)
    return mol_tree.succ, best_route, expanded + started_without_expansion


def predictor_expand_batch_fn(predictor, graph_fn: Callable[[str], object], condition_fn: Callable[[List[str]], object], topk: int = 50,
                              analysis_fn: Optional[Callable[[List[str]], list]] = None):
    """`expand_batch_fn` for `molstar_batched` on top of `GraphPredictor.sample_templates_batch`: `graph_fn(smiles)` -> PyG-like
    graph or None (the reference's `smiles_to_graph`, modeling_llamole.py:720-760), `condition_fn(list_of_smiles)` -> (len, 768)
    conditions (the LLM's retro query hidden states; None = the predictor's text_dropping row), `analysis_fn` -> per-molecule
    analysis tokens.  Returns dicts in the format of `one_step_reaction` (modeling_llamole.py:884-889)."""
    def expand(mols: List[str]):
        graphs = [graph_fn(m) for m in mols]
        ok = [i for i, g in enumerate(graphs) if g is not None]
        out: List[Optional[dict]] = [{"reactants": [], "scores": [], "templates": [], "analysis": []} for _ in mols]
        if not ok:
            return out
        c = condition_fn([mols[i] for i in ok])
        analyses = analysis_fn([mols[i] for i in ok]) if analysis_fn is not None else [[] for _ in ok]
        triples = predictor.sample_templates_batch([graphs[i] for i in ok], c, [mols[i] for i in ok], topk=topk)
        for i, (reactants, scores, templates), a in zip(ok, triples, analyses):
            out[i] = {"reactants": reactants, "scores": scores, "templates": templates, "analysis": a}
        return out

    return expand
