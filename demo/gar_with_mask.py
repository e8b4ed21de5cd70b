"""Single-region captioning CLI with the reference's flags (reference: demo/gar_with_mask.py:20-128),
running on the B200 engine.  Needs a GAR checkpoint directory (config.json + safetensors with the
reference's state-dict names) and its HF processor; `--synthetic` exercises the same code path with
random-init weights and a synthetic request when neither is available (no network in this image).

    torchrun --nproc-per-node 1 --master-addr 127.0.0.1 demo/gar_with_mask.py \
        --model_name_or_path <dir> --image_path assets/demo_image_1.png --mask_path assets/demo_mask_1.png
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from grasp_any_region_b200 import GARConfig, GARModel, config as presets  # noqa: E402

TORCH_DTYPE_MAP = dict(fp16=torch.float16, bf16=torch.bfloat16, fp32=torch.float32)


def parse_args():
    p = argparse.ArgumentParser(description="Inference demo of Grasp Any Region models (B200 engine).")
    p.add_argument("--model_name_or_path", help="HF model name or path", default="HaochenWang/GAR-8B")
    p.add_argument("--image_path", help="image path", required=False)
    p.add_argument("--mask_path", help="mask path", required=False)
    p.add_argument("--data_type", help="data dtype", type=str, choices=["fp16", "bf16", "fp32"], default="bf16")
    p.add_argument("--seed", type=int, default=0, help="Random seed for reproducible text generation")
    p.add_argument("--synthetic", default=None, choices=sorted(presets.PRESETS),
                   help="no checkpoint: random-init weights of this architecture + a synthetic request")
    p.add_argument("--max_new_tokens", type=int, default=1024)
    p.add_argument("--max_num_tiles", type=int, default=36, help="tile budget of the image processor (PLM default 36)")
    return p.parse_args()


def main():
    args = parse_args()
    if args.data_type != "bf16":
        raise SystemExit("the B200 engine computes in bf16 (--data_type bf16); fp16/fp32 are not built")
    torch.manual_seed(args.seed)
    torch.cuda.set_device(0)
    if "RANK" in os.environ and not torch.distributed.is_initialized():
        torch.distributed.init_process_group(backend="nccl")  # the reference does this too (:80-81)

    from transformers import GenerationConfig

    if args.synthetic:
        from grasp_any_region_b200.synth import make_request

        spec = presets.PRESETS[args.synthetic]()
        model = GARModel(GARConfig.from_spec(spec)).random_init_(args.seed).to("cuda:0").eval()
        if args.image_path and args.mask_path:
            # a real image + mask through the repo's own request builder (no hub processor / tokenizer needed:
            # the question's token ids are a fixed placeholder, so the printed ids are not a caption)
            from PIL import Image

            from grasp_any_region_b200.preprocess import build_request

            img = np.array(Image.open(args.image_path).convert("RGB"))
            mask = np.array(Image.open(args.mask_path).convert("L")).astype(bool)
            sample = build_request(spec, img, [mask], text_ids=list(range(1000, 1040)), device="cuda:0")
        else:
            sample = make_request(spec, seed=args.seed, n_masks=1)
        eos = pad = None
        decode = lambda ids: " ".join(str(int(t)) for t in ids)  # noqa: E731  (no tokenizer offline)
    else:
        from PIL import Image
        from transformers import AutoTokenizer

        from grasp_any_region_b200.request import single_region_request  # == SingleRegionCaptionDataset(...)[0]

        model = GARModel.from_pretrained(args.model_name_or_path, torch_dtype=torch.bfloat16,
                                         device_map="cuda:0").eval()
        tokenizer = AutoTokenizer.from_pretrained(args.model_name_or_path)
        img = np.array(Image.open(args.image_path).convert("RGB"))
        mask = np.array(Image.open(args.mask_path).convert("L")).astype(bool)
        # raw=True: the uint8 image and id map go to the GPU, resize / tiling / normalisation run there
        sample = single_region_request(model.spec, tokenizer, img, mask, prompt_token="<Prompt0>",
                                       max_num_tiles=args.max_num_tiles, raw=True)
        eos, pad = tokenizer.eos_token_id, tokenizer.pad_token_id
        decode = lambda ids: tokenizer.decode(ids, skip_special_tokens=True).strip()  # noqa: E731

    gen = GenerationConfig(max_new_tokens=args.max_new_tokens, do_sample=False, eos_token_id=eos, pad_token_id=pad)
    with torch.no_grad():
        if "image_u8" in sample:   # device-preprocessing request (B200 extension of the same call)
            model.configure_engine(max_new_tokens=args.max_new_tokens)
            ids = model.generate_batch([sample], max_new_tokens=args.max_new_tokens, eos_token_id=eos)[0]
        else:
            ids = model.generate(**sample, generation_config=gen, return_dict=True).sequences[0]
    print(decode(ids))


if __name__ == "__main__":
    main()
