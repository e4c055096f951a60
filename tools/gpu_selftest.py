"""Run every GPU parity check and print/save all measurements (does not stop at the first
failure).  Usage on the GPU box:  python tools/gpu_selftest.py [--quick] > gpurun_out/selftest.log"""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import gpu_checks as G
    quick = "--quick" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    jobs = [("layernorm", G.check_layernorm), ("linear_fp32", G.check_linear_fp32), ("linear_bf16", G.check_linear_bf16),
            ("linear_fp16", lambda: G.check_linear_bf16("fp16")),
            ("lnload_fp16", lambda: G.check_linear_ln_on_load("fp16")), ("lnload_bf16", lambda: G.check_linear_ln_on_load("bf16")),
            ("skinny_fp16", lambda: G.check_linear_skinny("fp16")), ("skinny_bf16", lambda: G.check_linear_skinny("bf16")),
            ("wattn_fp32", lambda: G.check_window_attention("fp32")), ("wattn_bf16", lambda: G.check_window_attention("bf16")),
            ("wattn_fp16", lambda: G.check_window_attention("fp16")),
            ("wattn_mma_fp16", lambda: G.check_window_attention("fp16", "mma")), ("wattn_mma_bf16", lambda: G.check_window_attention("bf16", "mma")),
            ("logsoftmax_topk", G.check_logsoftmax_topk), ("preprocess", G.check_preprocess),
            ("features", G.check_feature_extraction), ("ensemble", G.check_ensemble)]
    cases = ["tiny_e2e_peaky", "tiny_e2e_xavier", "feat_peaky_b5", "feat_xavier_b1"] + ([] if quick else ["full_e2e_xavier", "full_e2e_peaky"])
    for c in cases:
        jobs += [(f"enc:{c}", lambda c=c: G.check_encoder(c, "fp32")), (f"dec:{c}", lambda c=c: G.check_decoder(c, "fp32")),
                 (f"beam:{c}", lambda c=c: G.check_beam(c, "fp32"))]
    for c in ["tiny_e2e_peaky"] + ([] if quick else ["full_e2e_xavier"]):
        jobs += [(f"enc_bf16:{c}", lambda c=c: G.check_encoder(c, "bf16")), (f"beam_bf16:{c}", lambda c=c: G.check_beam(c, "bf16"))]
        jobs += [(f"enc_fp16:{c}", lambda c=c: G.check_encoder(c, "fp16")), (f"dec_fp16:{c}", lambda c=c: G.check_decoder(c, "fp16")),
                 (f"dec_bf16:{c}", lambda c=c: G.check_decoder(c, "bf16")), (f"beam_fp16:{c}", lambda c=c: G.check_beam(c, "fp16"))]
    jobs += [("dec_fp16:feat_peaky_b5", lambda: G.check_decoder("feat_peaky_b5", "fp16")),
             ("beam_fp16:feat_peaky_b5", lambda: G.check_beam("feat_peaky_b5", "fp16"))]
    # round 2
    jobs += [("graph_ptr", G.check_graph_pointer_independence), ("sampling", G.check_sampling),
             ("preprocess_batch", G.check_preprocess_batch), ("evaluate_loop", G.check_evaluate_model_loop),
             ("saturation", G.check_fp16_saturation), ("nvjpeg", G.check_nvjpeg_decode), ("early_exit", G.check_early_exit)]
    if not quick:
        jobs += [("e2e16_fp16:full_e2e_peaky", lambda: G.check_image_to_logits_16bit("full_e2e_peaky", "fp16")),
                 ("e2e16_fp16:full_e2e_xavier", lambda: G.check_image_to_logits_16bit("full_e2e_xavier", "fp16")),
                 ("e2e16_bf16:full_e2e_peaky", lambda: G.check_image_to_logits_16bit("full_e2e_peaky", "bf16")),
                 ("p3_288:enc", lambda: G.check_encoder("full_p3_288_n2", "fp32")), ("p3_288:dec", lambda: G.check_decoder("full_p3_288_n2", "fp32")),
                 ("p3_288:beam", lambda: G.check_beam("full_p3_288_n2", "fp32")),
                 ("e2e16_fp16:full_p3_288_n2", lambda: G.check_image_to_logits_16bit("full_p3_288_n2", "fp16")),
                 ("c2_b64:peaky_fp32", lambda: G.check_config2_batch64("c2_b64_peaky", "fp32")),
                 ("c2_b64:xavier_fp32", lambda: G.check_config2_batch64("c2_b64_xavier", "fp32")),
                 ("c2_b64:peaky_fp16", lambda: G.check_config2_batch64("c2_b64_peaky", "fp16")),
                 ("demo_c1", G.check_demo_known_answers)]
    if not quick:
        jobs += [("config3", G.check_config3_features_beam5), ("config4", G.check_config4_batch512_chunking),
                 ("caption_host", G.check_caption_host)]
    results, nbad = [], 0
    for name, fn in jobs:
        if only and not any(o in name for o in only):
            continue
        t0 = time.time()
        try:
            triples = fn()
            for (l, v, t) in triples:
                ok = v <= t
                nbad += (not ok)
                print(f"{'PASS' if ok else 'FAIL'}  {l}: {v:.3e} (tol {t:.1e})", flush=True)
                results.append(dict(check=name, label=l, value=v, tol=t, ok=bool(ok)))
        except Exception as ex:  # keep going: one GPU call should tell us as much as possible
            nbad += 1
            print(f"ERROR {name}: {ex!r}", flush=True)
            traceback.print_exc()
            results.append(dict(check=name, label="exception", value=None, tol=None, ok=False, error=repr(ex)))
            if "CUDA" in repr(ex) or "cuda" in repr(ex):
                print("CUDA error: stopping", flush=True)
                break
        print(f"      [{name} {time.time() - t0:.1f}s]", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = os.environ.get("SELFTEST_TAG", "all")
    with open(os.path.join(ROOT, "gpurun_out", f"selftest_{tag}.json"), "w") as f:
        json.dump(results, f, indent=1)
    print(f"SELFTEST {'OK' if nbad == 0 else 'FAILED'}: {nbad} failing of {len(results)}")
    return 0 if nbad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
