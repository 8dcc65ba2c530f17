"""String -> class plugin registry with honk2's contract.

Mirrors the interface of /root/reference/utils/class_registry.py:4-14 (``register_cls`` /
``find_cls``) over the dotted-key trie of /root/reference/utils/trie.py:4-32:
  * identifiers are split on '.', one trie level per token;
  * re-registering an identifier silently OVERWRITES the earlier class (trie.py:21) -- this is
    the plug-in hook that lets this package take over "model.ResNet" / "model.CNN";
  * ``find_cls`` of an unknown identifier returns ``default_value`` (None), never raises
    (trie.py:28-29).
"""


class _Node:
    __slots__ = ("value", "children")

    def __init__(self):
        self.value = None
        self.children = {}


class Registry:
    def __init__(self):
        self._root = _Node()

    def add(self, identifier, value):
        node = self._root
        for token in identifier.split("."):
            node = node.children.setdefault(token, _Node())
        node.value = value

    def get(self, identifier, default_value=None):
        node = self._root
        for token in identifier.split("."):
            node = node.children.get(token)
            if node is None:
                return default_value
        return node.value

    def identifiers(self, prefix=""):
        out = []

        def walk(node, path):
            if node.value is not None:
                out.append(".".join(path))
            for tok, child in sorted(node.children.items()):
                walk(child, path + [tok])

        walk(self._root, [])
        return [i for i in out if i.startswith(prefix)]


_REGISTRY = Registry()


def register_cls(identifier):
    def add_class(cls):
        _REGISTRY.add(identifier, cls)
        return cls

    return add_class


def find_cls(identifier, default_value=None):
    return _REGISTRY.get(identifier, default_value)


def install_into(reference_register_cls, prefixes=("model.",)):
    """Re-register this package's classes into ANOTHER registry (honk2's own
    ``utils.register_cls``) so that ``find_cls("model.ResNet")`` inside an unmodified honk2
    checkout resolves to the B200 implementation (overwrite semantics, trie.py:21).

    Only the ``model.*`` identifiers are installed by default: ``run/run_utils.py`` is imported by
    the training script too, and the device-resident metrics / loss of this package are inference
    tools.  ``prefixes=("model.", "metric.", "loss_fn.", "data_loader.")`` opts in to the rest (they
    keep the reference's ``get_type()`` / ``collect_metrics`` contract, metric/metric_utils.py:5-52).
    Returns the identifiers it installed."""
    done = []
    for ident in _REGISTRY.identifiers():
        if any(ident.startswith(p) for p in prefixes):
            reference_register_cls(ident)(_REGISTRY.get(ident))
            done.append(ident)
    return done
