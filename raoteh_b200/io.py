"""
Input formats on the caller side of the hot path, lowered straight to the batch tensors the
kernels take.

Readers with the reference's signatures and return values
(examples/p53/app_helper.py:23-184): `read_genetic_code`, `read_phylip`, `read_newick`
(own recursive-descent parser: the reference uses dendropy, which is not a dependency here;
the node numbering is the reference's -- leaves first in left-to-right order, then the
internal nodes in postorder, root last, :120-150), `read_disease_data`.

Lowering, replacing the per-column Python loops of examples/p53/p53.py:76-100 and
examples/p53/blink.py:229-270 (one dict of allowed states per column and one likelihood /
sampler call per column): `alignment_to_codes` turns the whole codon alignment into the
uint8 [n_leaves, n_columns] code matrix of `engine.Observations.from_leaf_codes`, and
`disease_to_tol_obs` turns the per-column human disease residues into the
uint8 [1, n_parts, n_columns] tolerance observations of `tmjp.ToleranceChains`.
"""
from __future__ import annotations

from collections import defaultdict

import networkx as nx
import numpy as np

MISSING = 255


# ---------------------------------------------------------------------------
# readers (examples/p53/app_helper.py)
# ---------------------------------------------------------------------------
def gen_paragraphs(lines):
    """examples/p53/app_helper.py:61-72"""
    para = []
    for line in lines:
        line = line.strip()
        if not line:
            if para:
                yield para
                para = []
        else:
            para.append(line)
    if para:
        yield para


def read_genetic_code(fin):
    """examples/p53/app_helper.py:163-184 -> list of (state, residue, codon), stops dropped."""
    genetic_code = []
    for line in fin:
        line = line.strip()
        if line:
            state, residue, codon = line.split()
            residue = residue.upper()
            if residue != 'STOP':
                genetic_code.append((int(state), residue, codon.upper()))
    return genetic_code


def read_phylip(fin, ntaxa=None, ncodons=None):
    """examples/p53/app_helper.py:75-99: yields (taxon name, codons).  The layout is the
    codeml one: a header line, then one paragraph per taxon (name line, then the sequence).
    A sequence written without blanks is cut into triplets.  `ntaxa` / `ncodons`, when given,
    are checked like the reference checks its 25 x 393 p53 alignment."""
    paras = list(gen_paragraphs(fin))
    if paras and all(tok.isdigit() for tok in paras[0][0].split()):   # "ntaxa nsites" header
        paras[0] = paras[0][1:]
        if not paras[0]:
            paras = paras[1:]
    # a paragraph is either a name line followed by codon lines, or "name  sequence" lines
    if ntaxa is not None and len(paras) != ntaxa:
        raise Exception('expected an alignment of %d taxa' % ntaxa)
    for para in paras:
        first = para[0].split()
        taxon_name = first[0]
        tokens = first[1:] + ' '.join(para[1:]).split()
        if len(tokens) == 1 or any(len(t) != 3 for t in tokens):
            seq = ''.join(tokens)
            if len(seq) % 3:
                raise Exception('sequence length of %s is not a multiple of three' % taxon_name)
            tokens = [seq[i:i + 3] for i in range(0, len(seq), 3)]
        if ncodons is not None and len(tokens) != ncodons:
            raise Exception('expected %d codons' % ncodons)
        yield taxon_name, tokens


class _Node(object):
    __slots__ = ('children', 'name', 'length')

    def __init__(self):
        self.children, self.name, self.length = [], None, None


def _parse_newick(text):
    text = text.strip()
    if not text.endswith(';'):
        raise ValueError('a newick string ends with a semicolon')
    pos = [0]

    def label():
        i = pos[0]
        if i < len(text) and text[i] in '"\'':
            q = text[i]
            j = text.index(q, i + 1)
            pos[0] = j + 1
            return text[i + 1:j]
        j = i
        while j < len(text) and text[j] not in ',():;[':
            j += 1
        pos[0] = j
        return text[i:j].strip()

    def skip_comment():
        while pos[0] < len(text) and text[pos[0]] == '[':
            pos[0] = text.index(']', pos[0]) + 1

    def node():
        nd = _Node()
        skip_comment()
        if text[pos[0]] == '(':
            pos[0] += 1
            while True:
                nd.children.append(node())
                skip_comment()
                if text[pos[0]] == ',':
                    pos[0] += 1
                    continue
                if text[pos[0]] == ')':
                    pos[0] += 1
                    break
                raise ValueError('malformed newick near position %d' % pos[0])
        skip_comment()
        name = label()
        nd.name = name or None
        skip_comment()
        if pos[0] < len(text) and text[pos[0]] == ':':
            pos[0] += 1
            nd.length = float(label())
        skip_comment()
        return nd
    root = node()
    if text[pos[0]] != ';':
        raise ValueError('malformed newick near position %d' % pos[0])
    return root


def read_newick(fin):
    """examples/p53/app_helper.py:102-150 -> (T, root_index, leaf_name_pairs): undirected
    weighted nx.Graph; leaves are numbered first (left to right), then the internal nodes in
    postorder, so the root is the last node."""
    text = fin.read() if hasattr(fin, 'read') else str(fin)
    root = _parse_newick(text)
    leaves, internal = [], []

    def walk(nd):
        for ch in nd.children:
            walk(ch)
        (internal if nd.children else leaves).append(nd)
    walk(root)
    ordered = leaves + internal
    index = dict((id(nd), i) for i, nd in enumerate(ordered))
    T = nx.Graph()
    T.add_nodes_from(range(len(ordered)))
    for nd in internal:
        for ch in nd.children:
            T.add_edge(index[id(nd)], index[id(ch)], weight=ch.length)
    leaf_name_pairs = [(i, str(nd.name)) for i, nd in enumerate(leaves)]
    return T, len(ordered) - 1, leaf_name_pairs


def read_disease_data(fin):
    """examples/p53/app_helper.py:23-41 -> dict column index -> set of disease residues
    (positions in the file are 1-based; insertions / deletions are skipped)."""
    column_to_disease_residues = defaultdict(set)
    for line in fin:
        line = line.strip()
        if not line:
            continue
        ntpos, codonpos, exon, wcodon, mcodon, wres, mres = line.split()
        wres, mres = wres.upper(), mres.upper()
        if wres == mres:
            raise Exception('synonymous disease: ' + line)
        if len(mcodon) != 3:
            if not ('INS' in mcodon or 'DEL' in mcodon):
                raise Exception('unrecognized mutant codon')
            continue
        column_to_disease_residues[int(codonpos) - 1].add(mres)
    return dict(column_to_disease_residues)


# ---------------------------------------------------------------------------
# lowering to batch tensors
# ---------------------------------------------------------------------------
def codon_state_maps(genetic_code):
    """(codon -> state, state -> residue, residue -> part, state -> part) with the parts
    numbered by sorted residue, as examples/p53/blink.py:97-111 builds them."""
    if [s for s, r, c in genetic_code] != list(range(len(genetic_code))):
        raise ValueError('the genetic code must list the states 0..n-1 in order '
                         '(examples/p53/p53.py:35-38)')
    codon_to_state = dict((c, s) for s, r, c in genetic_code)
    state_to_residue = dict((s, r) for s, r, c in genetic_code)
    residues = sorted(set(r for s, r, c in genetic_code))
    residue_to_part = dict((r, i) for i, r in enumerate(residues))
    state_to_part = dict((s, residue_to_part[r]) for s, r in state_to_residue.items())
    return codon_to_state, state_to_residue, residue_to_part, state_to_part


def alignment_to_codes(name_codons_list, leaf_name_pairs, sched, codon_to_state):
    """Whole alignment -> (codes uint8 [n_leaves, n_columns], leaf node indices of `sched`).
    Replaces the per-column dict building of examples/p53/p53.py:88-95.  Codons that are not in
    the table (gaps, stops, ambiguity codes) become MISSING = unobserved."""
    name_to_leaf = dict((name, leaf) for leaf, name in leaf_name_pairs)
    n_cols = len(name_codons_list[0][1])
    codes = np.full((len(name_codons_list), n_cols), MISSING, dtype=np.uint8)
    leaf_nodes = np.zeros(len(name_codons_list), dtype=np.int32)
    for i, (name, codons) in enumerate(name_codons_list):
        if len(codons) != n_cols:
            raise ValueError('sequences of unequal length')
        leaf_nodes[i] = sched.node_index[name_to_leaf[name]]
        for j, codon in enumerate(codons):
            codes[i, j] = codon_to_state.get(codon.upper(), MISSING)
    return codes, leaf_nodes


def disease_to_tol_obs(column_to_disease_residues, residue_to_part, n_parts, n_columns):
    """Per-column disease residues of the reference taxon -> tolerance observations
    uint8 [1, n_parts, n_columns] (bit0 = off allowed, bit1 = on allowed): a class holding a
    disease residue is observed OFF, every other class ON; a column without data is an
    observation that every amino acid is tolerated (examples/p53/blink.py:244-269)."""
    tol = np.full((1, n_parts, n_columns), 2, dtype=np.uint8)
    for col, residues in column_to_disease_residues.items():
        if 0 <= col < n_columns:
            for r in residues:
                tol[0, residue_to_part[r], col] = 1
    return tol


def mg94_from_genetic_code(A, C, G, T, kappa, omega, genetic_code, target_expected_rate=None,
                           target_expected_syn_rate=None):
    """The MG94 codon model in the state order of a genetic code table
    (examples/p53/create_mg94.py:22-135): rate to a codon one nucleotide away = frequency of
    the new nucleotide, x kappa for transitions, x omega for residue changes; stationary
    distribution = product of nucleotide frequencies; rescaled to an expected (synonymous)
    rate.  Returns (Q dense with diagonal, distn array, state -> residue, residue -> part)."""
    if (target_expected_rate is None) == (target_expected_syn_rate is None):
        raise ValueError('give exactly one of target_expected_rate and target_expected_syn_rate')
    states = [s for s, r, c in genetic_code]
    if states != list(range(len(genetic_code))):
        raise ValueError('the genetic code must list the states 0..n-1 in order')
    codon_to_state, state_to_residue, residue_to_part, _ = codon_state_maps(genetic_code)
    nt = dict(A=A, C=C, G=G, T=T)
    n = len(genetic_code)
    transitions = ('AG', 'GA', 'CT', 'TC')
    Q = np.zeros((n, n))
    syn = np.zeros((n, n), dtype=bool)
    for a, (sa, ra, ca) in enumerate(genetic_code):
        for b, (sb, rb, cb) in enumerate(genetic_code):
            diff = [(x, y) for x, y in zip(ca, cb) if x != y]
            if len(diff) != 1:
                continue
            x, y = diff[0]
            rate = nt[y] * (kappa if x + y in transitions else 1.0)
            if ra != rb:
                rate *= omega
            else:
                syn[a, b] = True
            Q[a, b] = rate
    distn = np.array([np.prod([nt[x] for x in c]) for s, r, c in genetic_code])
    distn /= distn.sum()
    flow = distn[:, None] * Q
    if target_expected_rate is not None:
        scale = target_expected_rate / flow.sum()
    else:
        scale = target_expected_syn_rate / flow[syn].sum()
    Q *= scale
    Q -= np.diag(Q.sum(axis=1))
    return Q, distn, state_to_residue, residue_to_part
