"""`Separator` facade: same constructor / load_model() / separate() surface as the reference orchestrator
(audio_separator/separator/separator.py:108-257, :830-933, :935-1045), restricted to what the accelerated path
needs.  Model download, ensembles, presets and file-level chunking are out of scope:
models are looked up in `model_file_dir` only.
"""
import importlib
import json
import logging
import os
import time

import torch

# UVR model_data_new.json entries for the BASELINE models (looked up by MD5 at runtime in the reference, separator.py:786-824)
KNOWN_MODEL_DATA = {
    "UVR-MDX-NET-Inst_HQ_3.onnx": {"compensate": 1.022, "mdx_dim_f_set": 3072, "mdx_dim_t_set": 8, "mdx_n_fft_scale_set": 6144, "primary_stem": "Instrumental"},
    "UVR-MDX-NET-Inst_HQ_5.onnx": {"compensate": 1.010, "mdx_dim_f_set": 2560, "mdx_dim_t_set": 8, "mdx_n_fft_scale_set": 5120, "primary_stem": "Instrumental"},
}


class Separator:
    def __init__(self, log_level=logging.INFO, log_formatter=None, model_file_dir="/tmp/audio-separator-models/", output_dir=None, output_format="WAV",
                 output_bitrate=None, normalization_threshold=0.9, amplification_threshold=0.0, output_single_stem=None, invert_using_spec=False,
                 sample_rate=44100, use_soundfile=False, use_autocast=False, use_directml=False, chunk_duration=None,
                 mdx_params=None, vr_params=None, demucs_params=None, mdxc_params=None, ensemble_algorithm=None, ensemble_weights=None,
                 ensemble_preset=None, info_only=False):
        self.logger = logging.getLogger(__name__)
        self.logger.setLevel(log_level)
        self.log_level = log_level
        if not self.logger.hasHandlers():
            h = logging.StreamHandler()
            h.setFormatter(log_formatter or logging.Formatter("%(asctime)s - %(levelname)s - %(module)s - %(message)s"))
            self.logger.addHandler(h)
        self.model_file_dir = os.environ.get("AUDIO_SEPARATOR_MODEL_DIR", model_file_dir)  # separator.py:167
        self.output_dir = output_dir or os.getcwd()
        self.output_format = output_format or "WAV"
        self.output_bitrate = output_bitrate
        if not 0 < normalization_threshold <= 1:
            raise ValueError("The normalization_threshold must be greater than 0 and less than or equal to 1.")
        if not 0 <= amplification_threshold <= 1:
            raise ValueError("The amplification_threshold must be greater than or equal to 0 and less than or equal to 1.")
        self.normalization_threshold, self.amplification_threshold = normalization_threshold, amplification_threshold
        self.output_single_stem, self.invert_using_spec = output_single_stem, invert_using_spec
        self.sample_rate = int(sample_rate)
        if self.sample_rate <= 0 or self.sample_rate > 12800000:
            raise ValueError(f"The sample rate setting is {self.sample_rate} but it must be a non-zero whole number.")
        self.use_soundfile, self.use_autocast = use_soundfile, use_autocast
        if ensemble_algorithm or ensemble_preset or chunk_duration:
            raise NotImplementedError("ensembles / file-level chunking are outside the B200 hot-path scope")
        self.arch_specific_params = {
            "MDX": {"hop_length": 1024, "segment_size": 256, "overlap": 0.25, "batch_size": 1, "enable_denoise": False, **(mdx_params or {})},
            "VR": {"batch_size": 1, "window_size": 512, "aggression": 5, "enable_tta": False, "enable_post_process": False, "post_process_threshold": 0.2, "high_end_process": False, **(vr_params or {})}, "Demucs": {"segment_size": "Default", "shifts": 2, "overlap": 0.25, "segments_enabled": True, **(demucs_params or {})}, "MDXC": {"segment_size": 256, "override_model_segment_size": False, "batch_size": 1, "overlap": 8, "pitch_shift": 0, **(mdxc_params or {})},
        }
        self.torch_device = self.torch_device_cpu = torch.device("cpu")
        self.torch_device_mps = None
        self.onnx_execution_provider = None
        self.model_instance = None
        self.model_is_uvr_vip = False
        self.model_friendly_name = None
        if not info_only:
            self.setup_accelerated_inferencing_device()

    def setup_accelerated_inferencing_device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: this build targets B200 (sm_100a) only and has no CPU / MPS / DirectML path")
        self.torch_device = torch.device("cuda", torch.cuda.current_device())
        self.logger.info(f"B200 engine on {torch.cuda.get_device_name(self.torch_device)}")

    def load_model_data(self, model_path):
        if model_path.lower().endswith(".th"):  # a single Demucs package carries its own constructor arguments (demucs/states.py:34-64)
            return {"b200_arch": "Demucs"}
        side = os.path.splitext(model_path)[0] + ".json"
        if os.path.exists(side):
            with open(side, encoding="utf-8") as f:
                return json.load(f)
        for ext in (".yaml", ".yml"):  # MDXC / Demucs models carry a YAML config next to the checkpoint (separator.py:758-777)
            if os.path.exists(os.path.splitext(model_path)[0] + ext):
                import yaml

                with open(os.path.splitext(model_path)[0] + ext, encoding="utf-8") as f:
                    return yaml.safe_load(f)
        name = os.path.basename(model_path)
        if name in KNOWN_MODEL_DATA:
            return dict(KNOWN_MODEL_DATA[name])
        raise ValueError(f"no model parameters for {name}: put a {os.path.basename(side)} (UVR model_data entry) next to the model file")

    def load_model(self, model_filename="UVR-MDX-NET-Inst_HQ_3.onnx"):
        if isinstance(model_filename, (list, tuple)):
            raise NotImplementedError("multi-model ensembles are outside the B200 hot-path scope")
        t0 = time.perf_counter()
        model_path = model_filename if os.path.isabs(model_filename) else os.path.join(self.model_file_dir, model_filename)
        if not os.path.isfile(model_path):
            raise FileNotFoundError(f"{model_path} not found (this build does not download models)")
        model_data = self.load_model_data(model_path)
        model_type = model_data.get("b200_arch") or ("VR" if "vr_model_param" in model_data else None) or ("MDXC" if "audio" in model_data and "model" in model_data else (
            "Demucs" if "models" in model_data and model_path.lower().endswith((".yaml", ".yml")) else ("MDX" if model_path.lower().endswith((".onnx", ".npz")) else None)))
        classes = {"MDX": "mdx_separator.MDXSeparator", "VR": "vr_separator.VRSeparator", "Demucs": "demucs_separator.DemucsSeparator", "MDXC": "mdxc_separator.MDXCSeparator"}
        if model_type not in classes:
            raise ValueError(f"Model type not supported (yet): {model_type}")
        common = {
            "logger": self.logger, "log_level": self.log_level, "torch_device": self.torch_device, "torch_device_cpu": self.torch_device_cpu,
            "torch_device_mps": self.torch_device_mps, "onnx_execution_provider": self.onnx_execution_provider,
            "model_name": os.path.splitext(os.path.basename(model_filename))[0], "model_path": model_path, "model_data": model_data,
            "output_format": self.output_format, "output_bitrate": self.output_bitrate, "output_dir": self.output_dir,
            "normalization_threshold": self.normalization_threshold, "amplification_threshold": self.amplification_threshold,
            "output_single_stem": self.output_single_stem, "invert_using_spec": self.invert_using_spec, "sample_rate": self.sample_rate,
            "use_soundfile": self.use_soundfile,
        }
        module_name, class_name = classes[model_type].split(".")
        module = importlib.import_module(f"{__package__}.architectures.{module_name}")
        self.model_instance = getattr(module, class_name)(common_config=common, arch_config=self.arch_specific_params[model_type])
        self.logger.info(f"Loading model completed in {time.perf_counter() - t0:.2f}s")

    def separate(self, audio_file_path, custom_output_names=None):
        if self.model_instance is None:
            raise ValueError("Initialization failed or model not loaded. Please load a model before attempting to separate.")
        paths = [audio_file_path] if isinstance(audio_file_path, str) else list(audio_file_path)
        outputs = []
        audio_ext = (".wav", ".flac", ".mp3", ".ogg", ".opus", ".m4a", ".aiff", ".ac3")  # separator.py:972
        for path in paths:
            if os.path.isdir(path):  # recursive, audio files only (separator.py:966-979)
                files = [os.path.join(root, f) for root, _, names in os.walk(path) for f in names if f.endswith(audio_ext)]
            else:
                files = [path]
            for f in files:
                try:
                    outputs.extend(self._separate_file(f, custom_output_names))
                except Exception as e:  # per-file errors are logged, not raised (separator.py:978-987)
                    self.logger.error(f"Failed to process file {f}: {e}")
        return outputs

    def _separate_file(self, audio_file_path, custom_output_names=None):
        t0 = time.perf_counter()
        out = self.model_instance.separate(audio_file_path, custom_output_names)
        self.model_instance.clear_gpu_cache()
        self.model_instance.clear_file_specific_paths()
        self.logger.info(f"Separation duration: {time.perf_counter() - t0:.2f}s")
        return out
