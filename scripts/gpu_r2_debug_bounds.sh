# the GPU suite under the debug build (device asserts + workspace guard bands), then the product suite
set +e
mkdir -p gpurun_out
echo "== debug build (SEQDIFF_DEBUG_BOUNDS=1: libseqdiff_b200_dbg.so) ==" > gpurun_out/debug_bounds_r02.log
for t in tests/test_decode_loss.py tests/test_structure_model.py tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_train.py; do
  SEQDIFF_DEBUG_BOUNDS=1 timeout 1500 python -m pytest $t -q -m gpu --timeout 900 -p no:cacheprovider 2>&1 | tail -2 >> gpurun_out/debug_bounds_r02.log
done
SEQDIFF_DEBUG_BOUNDS=1 python - >> gpurun_out/debug_bounds_r02.log 2>&1 <<'PY'
import ctypes, torch, seqdiff_b200 as sd
import bench as Bn
# negative control: break one band on purpose and see the check fire
dev = torch.device("cuda:0"); Bn.L = 128
common = dict(max_position_embeddings=128, intermediate_size=1024, num_hidden_layers=6, position_embedding_type="relative_key")
m = sd.ConditionalBertForDiffusionBase(sd.BertConfig(**common), sd.BertConfig(**common, is_decoder=True, add_cross_attention=True), 20).eval().to(dev)
b, x = Bn.synthetic_workload(4)
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
out = m(torch.full((4, 1), 7.0, device=dev), x.to(dev), d["ligand_angles"], d["ligand_attn_mask"], d["receptor_seq"], d["receptor_angles"], d["receptor_attn_mask"])
nb, nk = ctypes.c_int(), ctypes.c_int()
rc = sd.lib().seqdiff_debug_check_guards(ctypes.byref(nb), ctypes.byref(nk), None)
print("forward B=4: rc", rc, "bands", nb.value, "broken", nk.value, "lib", sd._cabi.LIB_PATH.split("/")[-1])
import os
os.environ["SEQDIFF_DEBUG_BREAK_GUARD"] = "1"   # negative control: the check overwrites 4 bytes of one band itself
rc = sd.lib().seqdiff_debug_check_guards(ctypes.byref(nb), ctypes.byref(nk), None)
print("negative control (one band deliberately overwritten): rc", rc, "bands", nb.value, "broken", nk.value, "->", sd.lib().seqdiff_last_error().decode())
PY
cat gpurun_out/debug_bounds_r02.log
exit 0
echo "== product build =="
for t in tests/test_gpu_ops.py tests/test_train.py tests/test_gpu_forward.py; do timeout 1500 python -m pytest $t -q -m gpu --timeout 900 -p no:cacheprovider 2>&1 | tail -2; done
