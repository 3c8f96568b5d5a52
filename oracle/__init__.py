"""CPU oracle for the VMM / Neural-EM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package
``multi_view_sort_b200`` never imports ``oracle`` and fails loudly when its CUDA
library is missing.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md §4), so the restatement in ``nem_oracle.py`` is pinned by executing
the reference's own code (``ref_import.py``, works only where
``/root/reference`` exists) on seeded inputs: ``tests/test_oracle_vs_reference``
asserts bit-equality on CPU, and ``make_golden.py`` freezes the reference's
outputs into ``tests/golden/`` so the pin travels to the GPU box.
"""
