"""Static check of the scripts that only run on the GPU box (bench.py, tools/, the package's host code): every
global name a function reads must be bound at module level or be a builtin.  The GPU-only legs of bench.py cannot
execute here, so a missing import inside one of them would otherwise first show up in the driver's run."""
import builtins
import pathlib
import symtable

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
FILES = sorted([ROOT / "bench.py", ROOT / "__graft_entry__.py", ROOT / "b200dit.py"]
               + list((ROOT / "omnihuman-1-hack_b200").glob("*.py")) + list((ROOT / "tools").glob("*.py"))
               + list((ROOT / "oracle").glob("*.py")))


def _unbound_globals(table, module_names, out):
    for child in table.get_children():
        for sym in child.get_symbols():
            name = sym.get_name()
            if sym.is_global() and sym.is_referenced() and not sym.is_assigned() and name not in module_names \
                    and not hasattr(builtins, name):
                out.append(f"{child.get_name()}:{name}")
            if sym.is_declared_global() and sym.is_assigned():
                module_names.add(name)
        _unbound_globals(child, module_names, out)


@pytest.mark.parametrize("path", FILES, ids=lambda p: str(p.relative_to(ROOT)))
def test_no_unbound_global_names(path):
    src = path.read_text()
    top = symtable.symtable(src, str(path), "exec")
    module_names = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    module_names |= {"__file__", "__name__", "__doc__", "__builtins__", "__spec__", "__package__", "__path__"}
    # names bound through `global x` inside functions
    def collect(t):
        for c in t.get_children():
            for s in c.get_symbols():
                if s.is_declared_global() and s.is_assigned():
                    module_names.add(s.get_name())
            collect(c)
    collect(top)
    bad = []
    _unbound_globals(top, module_names, bad)
    assert not bad, f"{path.name}: names read but never bound: {bad}"
