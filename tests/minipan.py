"""Minimal duck-typed stand-ins for the PPanGGOLiN data model (Pangenome / Organism / GeneFamily / Edge), exposing only
the accessors the NEM path reads (reference: pangenome.py:213-323, geneFamily.py:245-296,
edge.py:40-120, graph/makeGraph.py:114-136).  Test helper, not product code."""
from collections import defaultdict


class Organism:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"Org({self.name})"


class Edge:
    def __init__(self, source, target):
        self.source, self.target = source, target
        self._organisms = defaultdict(list)
        source._edges[target] = self
        target._edges[source] = self

    def get_organisms_dict(self):
        return self._organisms


class GeneFamily:
    def __init__(self, name):
        self.name = name
        self._orgs = {}
        self._edges = {}
        self.partition = ""

    @property
    def organisms(self):
        yield from self._orgs.keys()

    @property
    def edges(self):
        yield from self._edges.values()


class Pangenome:
    def __init__(self):
        self._fams = {}
        self._orgs = {}
        self._edge = {}
        self.status = defaultdict(lambda: "No")
        self.parameters = {}

    @property
    def gene_families(self):
        yield from self._fams.values()

    @property
    def organisms(self):
        yield from self._orgs.values()

    def get_gene_family(self, name):
        return self._fams[name]

    def add_edge(self, fam1, fam2, org, pair):
        key = frozenset([fam1, fam2])
        e = self._edge.get(key)
        if e is None:
            e = Edge(fam1, fam2)
            self._edge[key] = e
        e._organisms[org].append(pair)
        return e


def build_from_layouts(layouts, n_fam):
    """``layouts[g]`` = family ids along genome g's single circular contig (graph/makeGraph.py:114-136 semantics)."""
    pan = Pangenome()
    fams = [GeneFamily(f"fam{i}") for i in range(n_fam)]
    used = set()
    for g, lay in enumerate(layouts):
        org = Organism(f"org{g}")
        pan._orgs[org.name] = org
        prev = None
        first = None
        for pos, f in enumerate(lay):
            fam = fams[int(f)]
            fam._orgs[org] = True
            used.add(int(f))
            gene = (g, pos)
            if prev is not None:
                pan.add_edge(fam, prev[0], org, (gene, prev[1]))
            else:
                first = (fam, gene)
            prev = (fam, gene)
        if prev is not None and len(lay) > 0:
            pan.add_edge(first[0], prev[0], org, (first[1], prev[1]))
    for i in sorted(used):
        pan._fams[fams[i].name] = fams[i]
    for k in ("genomesAnnotated", "genesClustered", "neighborsGraph"):
        pan.status[k] = "Computed"
    pan.status["partitioned"] = "No"
    return pan


def build_from_contigs(genomes, n_fam):
    """``genomes[g]`` = list of contigs, each the family ids along a circular contig."""
    pan = Pangenome()
    fams = [GeneFamily(f"fam{i}") for i in range(n_fam)]
    used = set()
    for g, contigs in enumerate(genomes):
        org = Organism(f"org{g}")
        pan._orgs[org.name] = org
        for ci, lay in enumerate(contigs):
            prev = None
            first = None
            for pos, f in enumerate(lay):
                fam = fams[int(f)]
                fam._orgs[org] = True
                used.add(int(f))
                gene = (g, ci, pos)
                if prev is not None:
                    pan.add_edge(fam, prev[0], org, (gene, prev[1]))
                else:
                    first = (fam, gene)
                prev = (fam, gene)
            if prev is not None and len(lay) > 0:
                pan.add_edge(first[0], prev[0], org, (first[1], prev[1]))
    for i in sorted(used):
        pan._fams[fams[i].name] = fams[i]
    for k in ("genomesAnnotated", "genesClustered", "neighborsGraph"):
        pan.status[k] = "Computed"
    pan.status["partitioned"] = "No"
    return pan
