"""CPU oracle for the obs-rvc per-audio-window inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`obs-rvc_b200/`) may import,
link or execute anything in this package.  Allowed users: `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py`.

What it is: a CPU restatement (numpy f32 DSP + torch-CPU fp32 networks) of
`rvc::RvcInfer::infer` (/root/reference/rvc/src/rvc.rs:133-220) and the functions it calls.

Parity status of the restatement (see DESIGN.md "Oracle"):
  * DSP / glue (`dsp.py`): PINNED to the reference's own known-answer tests
    (rmvpe.rs:271-326 STFT table / pad_reflect / pad_constant, mel.rs:266-330 mel helpers and
    `testdata/mel_filters.npz`, rvc/src/tests/{input_wav,feats}.npy frame geometry).
  * Network bodies (`nets.py`): PARITY UNPINNED.  The reference executes opaque .onnx graphs
    through ONNX Runtime (`ort` 2.0.0-rc.2, Cargo.lock:843) and ships no weights and no
    numeric golden that can run without them; the bodies are restated from the public
    RVC / fairseq definitions and run with seeded synthetic weights.
  * kNN (`knn.py`): PARITY UNPINNED - the reference has only `// TODO: index search`
    (rvc.rs:159); upstream RVC semantics are restated.
"""
