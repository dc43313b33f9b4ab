"""VoxelTokenMemory -- drop-in for the reference's memory class (memory_2.py:38) on top of the CUDA path.

Same constructor, methods, attributes and return types as the reference for the spatial-memory hot path
(SURVEY.md section 8b); the simulator (`Env`), YOLO long-term memory and diffusion "imagination" are injected
collaborators, not part of this package.  Differences that are part of the contract:
  * tokens are committed on the device as frames are fused; `update_memory_dist_base()` is therefore a no-op flush;
  * the store keeps per cell the `cache_size` tokens with the smallest float32 r^2 (deterministic) where the
    reference evicts at random once a cell holds 10 tokens (memory_2.py:352);
  * `fuse_frames` is the batched entry point (many frames per call); `obs2voxeltoken` is the single-frame form.
"""
from __future__ import annotations

import gc
import json
import os

import numpy as np
import torch

from . import _lib
from .geometry import PoseChain, sim_cam_mat_with_fov
from .store import SpatialStore, StoreConfig
from .vit import ViTEngine, gaussian_pool


class _NullEnv:
    """Stand-in when no simulator is injected (query-only use, synthetic builds)."""
    sims = None
    agent = None
    plnner = None

    class _State:
        position = np.zeros(3)
    original_state = _State()

    def reset(self, *a, **k):
        return None


class VoxelTokenMemory:
    def __init__(self, args, memory_path=None, init_state=None, build_map=False, preload_dino=None, preload_yolo=None,
                 need_diffusion=True, env=None, cache_size=10, bev=False, max_slots=None, max_batch=64, colour=False):
        self.args = args
        self.device = "cuda"                                   # memory_2.py:41
        if preload_dino is None:
            raise _lib.BscnavError(
                "preload_dino is required: pass a bscnav ViTEngine (or any module with forward_features); there is no "
                "network here for torch.hub.load('facebookresearch/dinov2', ...) (memory_2.py:43)")
        self.dinov2 = preload_dino
        self.yolow = preload_yolo                               # long-term memory detector: out of scope, kept for callers
        if memory_path:
            self.memory_save_path = memory_path
        else:
            self.memory_save_path = os.path.join(getattr(args, "memory_path", "."), getattr(args, "scene_name", "scene"))
        self.Env = env if env is not None else _NullEnv()

        self.patch_h = self.patch_w = getattr(getattr(preload_dino, "cfg", None), "patch", 14)   # memory_2.py:80-81
        self.n_patch_w = args.query_width // self.patch_w
        self.n_patch_h = args.query_height // self.patch_h
        self._chain = PoseChain(args.base_forward_axis, args.base_left_axis, args.base_up_axis, args.base2cam_rot,
                                args.sensor_height)
        self.base_transform = self._chain.base_transform
        self.base2cam_tf = self._chain.base2cam_tf
        self.cs = args.cell_size
        self.gs = int(args.grid_size)
        self.depth_sample_rate = args.depth_sample_rate
        self.calib_mat = sim_cam_mat_with_fov(args.height, args.width, fov=90)      # memory_2.py:102
        self.min_depth = args.min_depth
        self.max_depth = args.max_depth
        self.token_dim = getattr(getattr(preload_dino, "cfg", None), "dim", 1024)  # memory_2.py:107
        self.cache_size = cache_size                                                # memory_2.py:111
        self.camera_height = args.sensor_height
        self.floor_height = args.floor_height
        self.map_height = args.map_height
        self.maxh = int(self.map_height / self.cs)
        self.minh = int(self.floor_height / self.cs)
        self.bev = bool(bev)
        self.colour = bool(colour)           # keep grid_rgb / weight / cv_map / max_height (memory_2.py:888-903)
        self._max_slots = max_slots
        self._max_batch = max_batch
        self.batch_query_min_slots = 32768
        self.long_memory_dict = []
        self.base_height = []
        self.floor_min_height = None
        self.floor_max_height = None
        self._store = None
        self._init_cache()
        self._pinned = {}
        self._copy_stream = None
        # unproject/scatter of batch i runs on a side stream next to the ViT of batch i+1 (different pipes: float64/XU and
        # HBM against the tensor cores); token buffers are double-buffered between the two
        self._fuse_stream = None
        self._tok_bufs = [None, None]
        self._tok_free = [None, None]
        self._tok_slot = 0
        self._last_fuse = None
        self._last_fuse_stream = None

    # ------------------------------------------------------------------ state
    def _init_cache(self):
        """Replaces memory_2.py:708-722: all buffers live in HBM (SpatialStore)."""
        a = self.args
        cfg = StoreConfig(height=a.height, width=a.width, grid_size=self.gs, cell_size=self.cs,
                          floor_height=self.floor_height, map_height=self.map_height, min_depth=self.min_depth,
                          max_depth=self.max_depth, n_patch_h=self.n_patch_h, n_patch_w=self.n_patch_w,
                          token_dim=self.token_dim, cache_size=self.cache_size, bev=self.bev, max_slots=self._max_slots,
                          calib_mat=self.calib_mat, colour=self.colour)
        if self._store is not None and self._store.cfg == cfg:
            self._join()
            self._store.reset()
        else:
            self._store = None
            gc.collect()
            torch.cuda.empty_cache()
            self._store = SpatialStore(cfg, device=self.device)
        self._chain.reset()
        self.iter_id = 0

    def _clear_memory(self):
        """memory_2.py:147-164."""
        self._join()
        self._store.reset()
        self._chain.reset()

    @property
    def inv_init_base_tf(self):
        return self._chain.inv_init_base_tf

    @property
    def max_id(self) -> int:
        self._join()
        return self._store.count()

    @property
    def grid_rgb_pos(self) -> np.ndarray:
        self._join()
        return self._store.positions()

    def _colour(self, i):
        if not self.colour:
            raise _lib.BscnavError("construct the memory with colour=True to keep grid_rgb / weight / cv_map / max_height")
        self._join()
        return self._store.colour_views()[i]

    @property
    def grid_rgb(self) -> np.ndarray:       # (max_id, 3) uint8, memory_2.py:896-898
        return self._colour(0)

    @property
    def weight(self) -> np.ndarray:         # (max_id,) float32, memory_2.py:899
        return self._colour(1)

    @property
    def cv_map(self) -> np.ndarray:         # (gs, gs, 3) uint8, memory_2.py:903
        return self._colour(2)

    @property
    def max_height(self) -> np.ndarray:     # (gs, gs) float64, memory_2.py:902
        return self._colour(3)

    @property
    def store(self) -> SpatialStore:
        self._join()               # whoever reaches for the store directly sees every scatter issued so far
        return self._store

    # ------------------------------------------------------------------ build
    def _get_patch_token(self, img: np.ndarray) -> torch.Tensor:
        """memory_2.py:732-742: (H,W,3|4) uint8 -> (n_patch_w, n_patch_h, D) device tokens."""
        t = torch.from_numpy(np.ascontiguousarray(img)).to(self.device).unsqueeze(0)
        return self._tokens_of(t)[0].reshape(self.n_patch_w, self.n_patch_h, -1)

    def _tokens_of(self, rgb_dev: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        if isinstance(self.dinov2, ViTEngine):
            return self.dinov2.forward_frames(rgb_dev, out=out)
        # any other module with the reference's plug-point signature: pre-process as the reference does
        x = rgb_dev[..., :3].permute(0, 3, 1, 2).float() / 255
        x = torch.nn.functional.interpolate(x, size=(self.args.query_height, self.args.query_width), mode="bilinear",
                                            align_corners=False, antialias=True)
        mean = torch.tensor([0.485, 0.456, 0.406], device=x.device).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225], device=x.device).view(1, 3, 1, 1)
        with torch.no_grad():
            return self.dinov2.forward_features((x - mean) / std)["x_norm_patchtokens"].float().contiguous()

    def _sample_indices(self, n_frames: int):
        """memory_2.py:746-749: a fresh permutation per frame, every depth_sample_rate-th entry.  rate == 1 keeps raster order."""
        if self.depth_sample_rate == 1:
            return None
        n = self.args.height * self.args.width
        idx = []
        for _ in range(n_frames):
            perm = np.arange(n)
            np.random.shuffle(perm)
            idx.append(perm[:: self.depth_sample_rate])
        return torch.from_numpy(np.stack(idx).astype(np.int32)).to(self.device)

    def _staging(self, name, slot, shape, dtype):
        """Pinned staging buffer `name` of input slot `slot` (two slots, alternating per call).  Before the host writes into
        it again, the host waits for the H2D copy that last read it (stream-side ordering alone would let a fast caller
        overwrite pixels the copy engine has not fetched yet)."""
        n = int(np.prod(shape))
        ent = self._pinned.get((name, slot))
        if ent is not None and ent[1] is not None:
            ent[1].synchronize()
        if ent is None or ent[0].numel() < n or ent[0].dtype != dtype:
            ent = [torch.empty(n, dtype=dtype, pin_memory=True), None]
            self._pinned[(name, slot)] = ent
        return ent[0][:n].view(*shape), ent

    def _inputs_to_device(self, rgb, depth):
        """Host frames -> device through a side copy stream and two device buffers, so the copy of batch i+1 overlaps
        the kernels of batch i.  Pinned torch tensors are copied directly (the caller keeps them unchanged until the copy
        has run, as with any non_blocking copy); numpy arrays go through per-slot pinned staging.

        Returns (rgb_dev, depth_dev, slot, late_depth).  The colour frames are on their way when this returns and the current
        stream waits for them.  For host depth, `depth_dev` is None and `late_depth()` -- to be called once the ViT has been
        enqueued -- stages and copies it and returns (depth_dev, event): only the scatter needs the depth, so at one frame per
        call its staging and its copy run behind the ViT instead of in front of it."""
        if torch.is_tensor(rgb) and rgb.is_cuda and torch.is_tensor(depth) and depth.is_cuda:
            return rgb.contiguous(), depth.contiguous(), None, None
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._in_bufs = [None, None]
            self._in_free = [None, None]
            self._in_slot = 0
        slot = self._in_slot = self._in_slot ^ 1
        staged_rgb = None
        if not torch.is_tensor(rgb):
            rgb = np.asarray(rgb)
            view, staged_rgb = self._staging("rgb", slot, rgb.shape, torch.uint8)
            view.numpy()[...] = rgb
            rgb = view
        if not torch.is_tensor(depth):
            depth = np.asarray(depth, dtype=np.float32)      # no copy for a float32 array; lists and other dtypes are converted here
        n_rgb = rgb.numel()
        n_dep = int(np.prod(depth.shape))
        cur = torch.cuda.current_stream()
        bufs = self._in_bufs[slot]
        if bufs is None or bufs[0].numel() < n_rgb or bufs[1].numel() < n_dep:
            # growing: the kernels that still read the old buffers finish first (they run on other streams than the one the
            # caching allocator would hand the memory back to); the new memory may be recycled from work queued on `cur`
            self._wait_events(self._in_free[slot])
            self._in_free[slot] = None
            bufs = self._in_bufs[slot] = (torch.empty(n_rgb, dtype=torch.uint8, device=self.device),
                                          torch.empty(n_dep, dtype=torch.float32, device=self.device))
            self._copy_stream.wait_stream(cur)
        rgb_d = bufs[0][:n_rgb].view(rgb.shape)
        if self._in_free[slot] is not None:
            for ev in self._in_free[slot]:
                self._copy_stream.wait_event(ev)                  # kernels that read these buffers two batches ago are done
        with torch.cuda.stream(self._copy_stream):
            rgb_d.copy_(rgb, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self._copy_stream)
        if staged_rgb is not None:
            staged_rgb[1] = copied
        cur.wait_event(copied)

        def copy_depth(depth):
            staged = None
            if not torch.is_tensor(depth):
                view, staged = self._staging("depth", slot, depth.shape, torch.float32)
                view.numpy()[...] = depth
                depth = view
            dep_d = bufs[1][:n_dep].view(depth.shape)
            with torch.cuda.stream(self._copy_stream):        # behind the colour copy, which already waited for the buffers' readers
                dep_d.copy_(depth, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            if staged is not None:
                staged[1] = ev
            return dep_d, ev

        if torch.is_tensor(depth) and depth.is_cuda:
            return rgb_d, depth.contiguous(), slot, None
        return rgb_d, None, slot, lambda: copy_depth(depth)

    @staticmethod
    def _wait_events(events):
        for ev in events or ():
            if ev is not None:
                ev.synchronize()

    def fuse_frames(self, rgb, depth, poses, sample_idx=None):
        """Batched build step: rgb (B,H,W,3|4) uint8, depth (B,H,W) float32, poses (B,7) -- pinned host tensors, numpy
        arrays or device tensors.  Returns without synchronising: copies run on a side stream, the ViT on the current one,
        unprojection + scatter on a second side stream (overlapping the next call's ViT)."""
        poses = np.asarray(poses, dtype=np.float64).reshape(-1, 7)
        tfs = self._chain.batch(poses)
        rgb, depth, slot, late_depth = self._inputs_to_device(rgb, depth)
        if sample_idx is None:
            sample_idx = self._sample_indices(rgb.shape[0])
        done = self.fuse_device(rgb, depth, tfs, sample_idx=sample_idx, late_depth=late_depth)
        if slot is not None:
            ev = torch.cuda.Event()
            ev.record()                                   # the ViT has read rgb; `done`: the scatter has read depth
            self._in_free[slot] = (ev, done)
        self.tf = self._chain.tf

    def begin_shard(self, origin_pose, first_frame: int):
        """One scene whose frames are split over ranks (dist.shard_frames, dist.fuse_scene_sharded): anchor the pose chain on the
        scene's FIRST pose (the reference takes the origin from its first observation, memory_2.py:844-851; a shard that starts
        later must not take its own) and number this rank's points from its shard's first frame, so that the token-store rule
        (smaller r^2, then the earlier point) decides across ranks as it does in a single build."""
        self._chain.reset()
        self._chain(np.asarray(origin_pose, dtype=np.float64).reshape(7))
        self._store.next_ordinal = int(first_frame) * self.args.height * self.args.width

    def fuse_device(self, rgb_dev, depth_dev, tfs, sample_idx=None, ord_base=None, late_depth=None):
        """ViT on the current stream, then unprojection + scatter on the side stream.  Returns the event that marks the
        end of the scatter.  Inputs are device tensors that stay valid until that event.  `late_depth` (fuse_frames with host
        frames): called after the ViT has been enqueued, returns (depth_dev, event the scatter waits for)."""
        main = torch.cuda.current_stream()
        if self._fuse_stream is None:
            self._fuse_stream = torch.cuda.Stream(device=self.device)
        B = rgb_dev.shape[0]
        k = self._tok_slot = self._tok_slot ^ 1
        buf = self._tok_bufs[k]
        shape = (B, self.n_patch_h * self.n_patch_w, self.token_dim)
        if buf is None or buf.shape[0] < B:
            self._wait_events([self._tok_free[k]])        # the side-stream scatter still reading the old buffer finishes first
            buf = self._tok_bufs[k] = torch.empty(shape, dtype=torch.float32, device=self.device)
            self._tok_free[k] = None
        if self._tok_free[k] is not None:
            main.wait_event(self._tok_free[k])            # the scatter that read this token buffer two calls ago is done
        engine = isinstance(self.dinov2, ViTEngine)
        tokens = self._tokens_of(rgb_dev, out=buf[:B]) if engine else self._tokens_of(rgb_dev)
        ready = torch.cuda.Event()
        ready.record(main)
        depth_ready = None
        if late_depth is not None:
            depth_dev, depth_ready = late_depth()
        side = self._fuse_stream if os.environ.get("BSCNAV_OVERLAP", "1") != "0" else main     # 0: everything on one stream (A/B)
        if side is not main:
            # tensors allocated on the main stream and read on the side stream: the allocator must not recycle them early
            for t in (depth_dev, sample_idx, rgb_dev if self.colour else None, None if engine else tokens,
                      tfs if torch.is_tensor(tfs) else None):
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(side)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            if depth_ready is not None:
                side.wait_event(depth_ready)
            self._store.fuse(depth_dev, tfs, tokens, sample_idx=sample_idx, ord_base=ord_base,
                             rgb=rgb_dev if self.colour else None)
            done = torch.cuda.Event()
            done.record(side)
        self._tok_free[k] = done
        self._last_fuse = done
        self._last_fuse_stream = side
        return done

    def _join(self):
        """Make the current stream wait for the last scatter (before reading the store)."""
        if self._last_fuse is not None:
            torch.cuda.current_stream().wait_event(self._last_fuse)
        self._last_fuse_stream = None          # what follows (merge, query, reset) runs on the current stream, ordered behind the scatter

    def max_id_async(self):
        """Non-blocking read of max_id: enqueues the device->host copy, returns a handle for max_id_wait().  The copy goes on the
        stream the last scatter ran on, behind it: the current stream does not wait for the scatter, so the next batch's ViT
        still overlaps it."""
        side = self._last_fuse_stream
        if side is None or self._last_fuse is None:
            return self._store.count_async()
        with torch.cuda.stream(side):
            handle = self._store.count_async()
        self._last_fuse = handle[1]           # a later _join() (reset, query, save) also waits for this copy
        return handle

    def max_id_wait(self, handle) -> int:
        return self._store.count_wait(handle)

    def obs2voxeltoken(self, obs, pose):
        """memory_2.py:842-903, one frame."""
        rgb = np.asarray(obs["rgb"])
        depth = np.asarray(obs["depth"], dtype=np.float32)
        self.fuse_frames(rgb[None], depth[None], np.asarray(pose)[None])

    def long_memory(self, obs):
        """YOLO-World long-term memory (memory_2.py:905-1025): detector is an out-of-scope collaborator."""
        return None

    def update_memory_dist_base(self):
        """memory_2.py:326-358: the reference flushes its 50 000-row host buffer to HDF5 here.  Tokens are already in the
        device store; this only surfaces pending errors (slot capacity, fp16 range of the feature extractor)."""
        self._store.count()
        if isinstance(self.dinov2, ViTEngine):
            self.dinov2.check_range()

    def update_memory_surp_base(self):
        """memory_2.py:364-536, the surprise-based store rule (accept a token when its cosine distance to the 26 neighbouring voxels'
        tokens exceeds a threshold, merge near-duplicates by union-find, average the groups).  No driver or method of the reference
        calls it (the build loop flushes through update_memory_dist_base, :881 and :1135), and every decision reads the groups the previous
        tokens wrote, so it is sequential per token by construction.  Not built: the distance-based rule is the store contract
        (SURVEY 8a)."""
        raise NotImplementedError("update_memory_surp_base has no call site in the reference (memory_2.py:364-536 is dead code); "
                                  "this memory keeps the distance-based store rule of update_memory_dist_base")

    def initial_memory(self):
        """memory_2.py:298-309: never overwrite an existing directory."""
        count = 1
        while os.path.exists(self.memory_save_path):
            self.memory_save_path = self.memory_save_path.split("_")[0]
            self.memory_save_path = f"{self.memory_save_path}_{count}"
            count += 1
        os.makedirs(self.memory_save_path)
        self.feat_path = self.memory_save_path + "/feat.h5df"

    # ------------------------------------------------------------------ query
    def imaginary(self, text_prompts, vis=False):
        raise NotImplementedError("text -> image generation (SD3.5, memory_2.py:258-276) is an injected collaborator: "
                                  "assign memory.imaginary = callable returning an object with .images")

    def _query_vector(self, prompt) -> torch.Tensor:
        if torch.is_tensor(prompt) or isinstance(prompt, np.ndarray):           # direct query-vector injection
            return torch.as_tensor(prompt, dtype=torch.float32).reshape(-1, self.token_dim).to(self.device)
        from torchvision import transforms
        tf = transforms.Compose([                                               # memory_2.py:66-70 (host, PIL)
            transforms.Resize((self.args.query_height, self.args.query_width)),
            transforms.ToTensor(),
            transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
        ])
        if isinstance(prompt, str):
            gen = self.imaginary(prompt)
            imgs = [gen.images[i] for i in range(self.args.imagenary_num)]
        else:
            imgs = [prompt]
        x = torch.stack([tf(im) for im in imgs]).to(self.device)                # memory_2.py:577-580
        with torch.no_grad():
            tokens = self.dinov2.forward_features(x)["x_norm_patchtokens"].float().contiguous()
        return gaussian_pool(tokens)                                            # memory_2.py:592-608

    def voxel_localized(self, prompt, K=100, batch_size=300, region_radius=np.inf, curr_grid=None):
        """memory_2.py:563-671.  Returns (best (1,3), top-K positions (K',3) int64, top-K similarities (K',) float64).
        `batch_size` is accepted and ignored (the reference overwrites it at :593)."""
        self._join()
        q = self._query_vector(prompt)
        floor = None
        if getattr(self.args, "load_single_floor", False) and self.floor_min_height is not None:
            floor = (self.floor_min_height, self.floor_max_height)
        # large stores: the tensor-core candidate + exact re-score path reads the fp16 shadow (half the bytes of the fp32 scan) and
        # returns bit-identical results; small stores: the plain exact scan (no shadow to maintain)
        if self._store.count() >= self.batch_query_min_slots:
            sims, pos, cnt = self._store.query_batch(q[:1], k=K, region_radius=region_radius, curr_grid=curr_grid, floor_range=floor)
        else:
            sims, pos, cnt = self._store.query(q[:1], k=K, region_radius=region_radius, curr_grid=curr_grid, floor_range=floor)
        n = int(cnt[0])
        top_pos = pos[0, :n].cpu().numpy().astype(np.int64)
        top_sim = sims[0, :n].cpu().numpy().astype(np.float64)
        return top_pos[0:1].copy(), top_pos, top_sim

    def voxel_localized_batch(self, queries, K=100, region_radius=np.inf, curr_grid=None):
        """Batched form: queries (Q, D) -> device tensors (sims (Q,K), positions (Q,K,3), counts (Q,)); the same values as Q
        calls of voxel_localized, computed by the tensor-core candidate + exact re-scoring path (SpatialStore.query_batch)."""
        self._join()
        floor = None
        if getattr(self.args, "load_single_floor", False) and self.floor_min_height is not None:
            floor = (self.floor_min_height, self.floor_max_height)
        return self._store.query_batch(self._query_vector(queries), k=K, region_radius=region_radius, curr_grid=curr_grid,
                                       floor_range=floor)

    def long_memory_filter(self):
        """memory_2.py:686-704."""
        if getattr(self.args, "load_single_floor", False) and self.floor_min_height is not None:
            return [o for o in self.long_memory_dict if self.floor_min_height <= o["loc"][2] <= self.floor_max_height]
        return self.long_memory_dict

    # ------------------------------------------------------------------ small helpers of the reference class
    def get_total_token_count(self) -> int:
        """memory_2.py:312-323 (sum of the HDF5 groups' feature rows): stored token rows, counted on the device."""
        self._join()
        st = self._store
        n = st.count() * st.g.cache
        total = int((st.row_key[:n] != -1).sum().item()) if n else 0
        print("total_tokens:", total)
        return total

    def read_memory(self):
        """memory_2.py:539-540: a stub in the reference as well."""

    def _reinit_cache(self):
        """memory_2.py:724-729 re-zeroes the 50 000-token host staging buffers after a flush; tokens go straight into the
        device store here, so there is nothing to re-initialise."""

    def _out_of_range(self, row, col, height):
        """memory_2.py:755-756."""
        return col >= self.gs or row >= self.gs or height >= self.maxh or col < 0 or row < 0 or height < self.minh

    def grid2loc_2d(self, x, y):
        """memory_2.py:1148-1158: BEV cell (row, col) -> world (x, z, y) at the initial height."""
        initial_x, initial_z, initial_y = self.Env.original_state.position
        return np.array([initial_x + (y - self.gs // 2) * self.cs, initial_z, initial_y + (x - self.gs // 2) * self.cs])

    def loc2grid_2d(self, x_base, y_base):
        """memory_2.py:1160-1163."""
        return int(self.gs / 2 - int(x_base / self.cs)), int(self.gs / 2 - int(y_base / self.cs))

    def is_unknown(self, x: int, y: int) -> bool:
        """memory_2.py:1165-1166 (needs colour=True: the BEV colour map is the occupancy evidence)."""
        return bool(self.cv_map[x, y].sum() == 0)

    def is_known(self, x: int, y: int) -> bool:
        return not self.is_unknown(x, y)

    def in_bounds(self, x: int, y: int) -> bool:
        return 0 <= x < self.gs and 0 <= y < self.gs

    def is_navigabale(self, x: int, y: int) -> bool:
        """memory_2.py:1171-1172 (spelling kept): delegated to the injected simulator's path finder."""
        return self.Env.plnner.pathfinder.is_navigable(self.grid2loc_2d(x, y))

    # ------------------------------------------------------------------ exploration drivers (simulator injected)
    def excute(self, obs, actions):
        """memory_2.py:1086-1101 without the cv2 preview."""
        for action in actions:
            if action != "stop":
                obs = self.Env.sims.step(action)
                st = self.Env.agent.get_state()
                pos, rot = st.position, st.rotation
                pose = np.array([pos[0], pos[1], pos[2], rot.x, rot.y, rot.z, rot.w])
                self.obs2voxeltoken(obs, pose)
                self.long_memory(obs)
        return obs

    def exploring_create_memory(self):
        """memory_2.py:1104-1145: random navigable sub-goals, look-around, save."""
        self.initial_memory()
        obs = self.Env.sims.get_sensor_observations(0)
        self.init_height = self.Env.agent.get_state().position[1]
        pf = self.Env.plnner.pathfinder
        for _ in range(self.args.random_move_num):
            subgoal = pf.get_random_navigable_point()
            island_begin = pf.get_island(self.Env.agent.get_state().position)
            while (not pf.is_navigable(subgoal)) or (pf.get_island(subgoal) != island_begin):
                subgoal = pf.get_random_navigable_point()
            try:
                path, _goal = self.Env.move2point(subgoal)
                obs = self.excute(obs, path)
                self.base_height.append(self.Env.agent.get_state().position[1])
                obs = self.excute(obs, ["turn_left"] * int(360 / self.args.turn_left))
            except Exception as e:   # noqa: BLE001  (the reference swallows motion failures the same way, :1126-1128)
                print(f"move failed: {e}")
                continue
        self.update_memory_dist_base()
        self.save_memory()

    def create_memory(self):
        raise NotImplementedError("keyboard-driven exploration (memory_2.py:1027-1083) needs an interactive simulator window")

    # ------------------------------------------------------------------ persistence (npy set of memory_2.py:1136-1145)
    def save_memory(self, path=None):
        from . import io
        self._join()
        io.save_memory(self, path or self.memory_save_path)

    def load_memory(self, init_state=None, build_map=False):
        """memory_2.py:166-256."""
        from . import io
        self._clear_memory()
        if hasattr(self.Env, "reset"):
            self.Env.reset(self.args, init_state=init_state, build_map=build_map)
        self.memory_save_path = self.args.load_memory_path
        if not build_map:
            io.load_memory(self, self.memory_save_path)


Memory = VoxelTokenMemory   # the name BASELINE.json's north_star uses
