"""Multi-region relationship CLI with the reference's flags (reference: demo/gar_relationship.py:26-143),
running on the B200 engine.  `--mask_paths` is a python list literal of mask files, one region each
(<Prompt0>, <Prompt1>, ...); the question gets the reference's multiple-choice suffix (:113-114).
`--synthetic <preset>` runs the same path with random-init weights and a synthetic multi-mask request."""
import argparse
import ast
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from grasp_any_region_b200 import GARConfig, GARModel, config as presets  # noqa: E402


def parse_args():
    p = argparse.ArgumentParser(description="Relationship demo of Grasp Any Region models (B200 engine).")
    p.add_argument("--model_name_or_path", help="HF model name or path", default="HaochenWang/GAR-8B")
    p.add_argument("--image_path", help="image path", required=False)
    p.add_argument("--mask_paths", help="python list literal of mask paths", type=str, default="[]")
    p.add_argument("--question_str", help="question with <Prompt{i}> references", type=str, default="")
    p.add_argument("--data_type", type=str, choices=["fp16", "bf16", "fp32"], default="bf16")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--synthetic", default=None, choices=sorted(presets.PRESETS))
    p.add_argument("--num_masks", type=int, default=3, help="--synthetic only")
    p.add_argument("--max_new_tokens", type=int, default=1024)
    p.add_argument("--max_num_tiles", type=int, default=36, help="tile budget of the image processor (PLM default 36)")
    p.add_argument("--no_reference_quirks", action="store_true",
                   help="give every region its own box (the reference's builder gives all regions the LAST mask's box)")
    return p.parse_args()


def main():
    args = parse_args()
    if args.data_type != "bf16":
        raise SystemExit("the B200 engine computes in bf16 (--data_type bf16)")
    torch.manual_seed(args.seed)
    torch.cuda.set_device(0)
    if "RANK" in os.environ and not torch.distributed.is_initialized():
        torch.distributed.init_process_group(backend="nccl")
    from transformers import GenerationConfig

    if args.synthetic:
        from grasp_any_region_b200.synth import make_request

        spec = presets.PRESETS[args.synthetic]()
        model = GARModel(GARConfig.from_spec(spec)).random_init_(args.seed).to("cuda:0").eval()
        sample = make_request(spec, seed=args.seed, n_masks=args.num_masks)
        eos = pad = None
        decode = lambda ids: " ".join(str(int(t)) for t in ids)  # noqa: E731
    else:
        from PIL import Image
        from transformers import AutoTokenizer

        from grasp_any_region_b200.request import multi_region_request  # == MultiRegionDataset(...)[0], quirks included

        model = GARModel.from_pretrained(args.model_name_or_path, torch_dtype=torch.bfloat16,
                                         device_map="cuda:0").eval()
        tokenizer = AutoTokenizer.from_pretrained(args.model_name_or_path)
        img = np.array(Image.open(args.image_path).convert("RGB"))
        masks = [np.array(Image.open(m).convert("L")).astype(bool) for m in ast.literal_eval(args.mask_paths)]
        question = args.question_str + "\nAnswer with the correct option's letter directly."
        sample = multi_region_request(model.spec, tokenizer, img, masks, question, max_num_tiles=args.max_num_tiles,
                                      device="cuda:0", reference_quirks=not args.no_reference_quirks)
        eos, pad = tokenizer.eos_token_id, tokenizer.pad_token_id
        decode = lambda ids: tokenizer.decode(ids, skip_special_tokens=True).strip()  # noqa: E731

    with torch.no_grad():
        out = model.generate(**sample, generation_config=GenerationConfig(
            max_new_tokens=args.max_new_tokens, do_sample=False, eos_token_id=eos, pad_token_id=pad),
            return_dict=True)
    print(decode(out.sequences[0]))


if __name__ == "__main__":
    main()
