"""TEST INFRASTRUCTURE ONLY.

CPU restatement of castorini/howl's data-parallel hot path (audio frontend -> res8
forward/backward -> AdamW).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package.
The product path (``howl_b200``) never imports it and has no CPU fallback.
"""
