"""HTTP surface of the reference's api/app.py (routes, multipart fields, validation ranges, status codes:
reference api/app.py:121-223), served by a resident FrameInterpolator instead of one subprocess + checkpoint load per
request (reference api/app.py:82-101). Concurrent uploads are micro-batched: a worker thread gathers the frame pairs
that arrive within FI_BATCH_WAIT_MS (default 2 ms, at most FI_MAX_BATCH = 16) and runs them as one forward."""
import asyncio
import concurrent.futures
import os
import queue
import threading
import uuid

import cv2
import numpy as np
from fastapi import FastAPI, File, Form, HTTPException, UploadFile
from fastapi.middleware.cors import CORSMiddleware
from fastapi.responses import FileResponse
from starlette.background import BackgroundTask

MODEL_PATH = os.environ.get("FI_MODEL_PATH", "best_model.pth")
OUTPUT_DIR = os.environ.get("FI_OUTPUT_DIR", "temp_outputs")
TARGET_SIZE = (256, 256)  # reference preprocess_image default (model/inference.py:11)

app = FastAPI(title="Frame Interpolation API", description="B200-native UNet frame interpolation", version="1.0.0")
app.add_middleware(CORSMiddleware, allow_origins=["*"], allow_credentials=True, allow_methods=["*"],
                   allow_headers=["*"])

_worker, _worker_lock = None, threading.Lock()    # _worker_lock: held by the batcher thread around GPU work
_batcher, _batcher_lock = None, threading.Lock()  # _batcher_lock: creation only, never held during a forward
_written = set()                                   # result files this process created and has not deleted yet


def get_worker():
    """One FrameInterpolator (weights + activation arena on the GPU) for the whole process; the handle is not
    re-entrant, so requests are serialised on it."""
    global _worker
    if _worker is None:
        from model.inference import FrameInterpolator
        _worker = FrameInterpolator(MODEL_PATH, os.environ.get("FI_DEVICE", "cuda"))
    return _worker


class _Batcher:
    """Single consumer of the GPU handle: drains the pending frame pairs into one batched forward."""

    def __init__(self, max_batch, max_wait_s):
        self.max_batch, self.max_wait_s = max_batch, max_wait_s
        self.pending = queue.Queue()
        self.batch_sizes = []            # sizes of the forwards run so far (observability / tests)
        threading.Thread(target=self._run, daemon=True).start()

    def submit(self, a, b):
        fut = concurrent.futures.Future()
        self.pending.put((a, b, fut))
        return fut

    def _run(self):
        import time
        while True:
            items = [self.pending.get()]
            deadline = time.monotonic() + self.max_wait_s
            while len(items) < self.max_batch:
                try:
                    items.append(self.pending.get(timeout=max(deadline - time.monotonic(), 0.0)))
                except queue.Empty:
                    break
            try:
                with _worker_lock:
                    mids = get_worker()._forward_pairs([i[0] for i in items], [i[1] for i in items])
                self.batch_sizes.append(len(items))
                for (_, _, fut), mid in zip(items, mids):
                    fut.set_result(mid)
            except Exception as e:  # noqa: BLE001  (every waiting request gets the error)
                for _, _, fut in items:
                    fut.set_exception(e)


def get_batcher():
    global _batcher
    if _batcher is not None:     # fast path: the event loop never waits behind a running forward
        return _batcher
    with _batcher_lock:
        if _batcher is None:
            _batcher = _Batcher(int(os.environ.get("FI_MAX_BATCH", "16")),
                                float(os.environ.get("FI_BATCH_WAIT_MS", "2")) / 1e3)
    return _batcher


def _decode(upload: UploadFile, data: bytes):
    if not upload.content_type or not upload.content_type.startswith("image/"):
        raise HTTPException(status_code=400, detail=f"{upload.filename} must be an image file")
    img = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise HTTPException(status_code=400, detail=f"could not decode {upload.filename}")
    return cv2.resize(img, TARGET_SIZE)


@app.post("/interpolate")
async def interpolate(frame1: UploadFile = File(...), frame2: UploadFile = File(...),
                      num_intermediate: int = Form(3), fps: int = Form(30)):
    if not 1 <= num_intermediate <= 10:
        raise HTTPException(status_code=400, detail="num_intermediate must be between 1 and 10")
    if not 10 <= fps <= 60:
        raise HTTPException(status_code=400, detail="fps must be between 10 and 60")
    a = _decode(frame1, await frame1.read())
    b = _decode(frame2, await frame2.read())
    try:
        from model.inference import save_frames_as_video
        mid = await asyncio.wrap_future(get_batcher().submit(a, b))
        os.makedirs(OUTPUT_DIR, exist_ok=True)
        path = os.path.join(OUTPUT_DIR, f"{uuid.uuid4().hex}.mp4")
        _written.add(path)
        # same frame list as inference.py:262-283; encoding runs off the event loop
        await asyncio.to_thread(save_frames_as_video, [a] + [mid] * num_intermediate + [b], path, fps)
        # the file is deleted as soon as the response has been sent (the reference reuses one video.mp4 per request)
        return FileResponse(path, media_type="video/mp4", filename="interpolated_video.mp4",
                            background=BackgroundTask(_discard, path))
    except HTTPException:
        raise
    except Exception as e:
        raise HTTPException(status_code=500, detail=f"Inference failed: {e}")


@app.get("/")
async def root():
    return {"message": "Frame Interpolation API", "version": "1.0.0",
            "endpoints": {"POST /interpolate": "Upload two frames and get an interpolated video",
                          "GET /health": "Health check"}}


@app.get("/health")
async def health():
    return {"status": "healthy", "model_exists": os.path.exists(MODEL_PATH), "model_path": MODEL_PATH}


def _discard(path):
    _written.discard(path)
    try:
        os.remove(path)
    except OSError:
        pass


@app.on_event("shutdown")
async def _cleanup():
    for path in list(_written):      # only what this process wrote: FI_OUTPUT_DIR may be a shared directory
        _discard(path)
    try:
        os.rmdir(OUTPUT_DIR)         # succeeds only when nothing else lives there
    except OSError:
        pass
