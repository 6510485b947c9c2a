"""Minimal stand-ins with the attribute surface of the reference's object model
(BGClib/BGClib.py: BGCCollection :750, BGC :882, BGCProtein :1828, BGCDomain
:3222) for tests that must run where /root/reference does not exist (the GPU
box).  tests/test_reference_objects.py runs the same checks on the real classes
when the reference is present."""


class FakeDomain:
    def __init__(self, protein, ID, alias, ali_from, ali_to):
        self.protein, self.ID, self.alias = protein, ID, alias
        self.ali_from, self.ali_to = ali_from, ali_to

    def get_sequence(self):                      # BGClib/BGClib.py:3240-3241
        return self.protein.sequence[self.ali_from:self.ali_to]


class FakeProtein:
    def __init__(self, identifier, sequence, parent_cluster_id=""):
        self.identifier = identifier
        self.parent_cluster_id = parent_cluster_id
        self.protein_type = ""
        self.role = "unknown"
        self.domain_list = []
        self.sequence = sequence

    @property
    def sequence(self):
        return self._sequence

    @sequence.setter
    def sequence(self, seq):                     # BGClib/BGClib.py:1903-1907
        self._sequence = seq.replace("\n", "").strip().replace(" ", "").upper()
        self.length = len(seq)


class FakeBGC:
    def __init__(self, identifier):
        self.identifier = identifier
        self.protein_list = []
        self.proteins = {}
        self.CBPcontent = {}


class FakeCollection:
    def __init__(self):
        self.bgcs = {}
        self.name = ""

    @classmethod
    def from_file(cls, collection_file):         # BGClib/BGClib.py:758-777
        import pickle
        assert collection_file.is_file()
        with open(collection_file, "rb") as dc:
            data = pickle.load(dc)
        assert isinstance(data, cls)
        return data


class FakeProteinCollection:
    def __init__(self):
        self.proteins = {}


def build_collection(spec, cbp=True):
    """spec: {bgc_id: [(protein_id, full_sequence, [(dom_ID, alias, from, to), ...]), ...]}"""
    col = FakeCollection()
    for bid, prots in spec.items():
        bgc = FakeBGC(bid)
        for pid, seq, doms in prots:
            p = FakeProtein(pid, seq, bid)
            for (did, alias, a0, a1) in doms:
                p.domain_list.append(FakeDomain(p, did, alias, a0, a1))
            bgc.protein_list.append(p)
            bgc.proteins[pid] = p
            if cbp:
                bgc.CBPcontent.setdefault("cbp", []).append(p)
        col.bgcs[bid] = bgc
    return col
