// irp_addon.cc — Node N-API addon: Buffer marshalling over the C ABI of libirp_b200.so.
//
// SOURCE ONLY / UNVERIFIED: this image has no Node toolchain (no node, npm, node_api.h), so this
// file has never been compiled.  It is deliberately logic-free: every decision lives behind
// include/irp.h, which IS exercised (from Python via ctypes, tests/ -m gpu).  Build, when Node is
// available:  node-gyp / cmake-js with  -I../include -L../image-restoration-platform_b200 -lirp_b200
//
// JS surface (used by addon/classifier.js and addon/imagePreprocess.js):
//   createContext(device:number) -> external
//   analyzeFile(ctx, file:Buffer) -> Promise<Float64Array(7)>   (JPEG bytes, baseline or progressive, decoded on the device)
//   transcodeFile(ctx, file:Buffer, orientation, quality) -> Promise<{scores:Float64Array(7), file:Buffer, width, height, channels}>
//       analyze() + preprocessImage() of one upload with files on both sides (irp_transcode_jpeg_batch)
//       `quality` carries the flags of include/irp.h: IRP_JPEG_OPTIMIZE, IRP_JPEG_ICC(id) (IRP_ICC_SRGB = the generated sRGB profile)
//   setOutputIcc(ctx, profile:Buffer|null)   the context DEFAULT profile (files encoded without IRP_JPEG_ICC bits); set once at start-up
//   analyzeRaw(ctx, pixels:Buffer, width, height, channels, isJpeg:boolean) -> Promise<Float64Array(7)>
//   preprocessRaw(ctx, pixels:Buffer, width, height, channels, orientation) -> Promise<{data:Buffer,width,height,channels}>
// Work runs on the libuv pool through napi_create_async_work, so the event loop never blocks
// (SURVEY.md §8b "Callers & threading"); the input Buffer is pinned alive with a napi_ref.
#include <node_api.h>

#include <cstdlib>
#include <cstring>
#include <string>

#include "../include/irp.h"

namespace {

struct Job {
  napi_async_work work = nullptr;
  napi_deferred deferred = nullptr;
  napi_ref input_ref = nullptr;
  irp_ctx* ctx = nullptr;
  irp_image_desc desc{};
  bool preprocess = false;
  bool file = false;            // the input Buffer is a JPEG FILE, decoded on the device (irp_submit_jpeg)
  irp_jpeg_desc jpeg{};
  bool transcode = false;       // file in, scores + preprocessed file out
  int quality = 85;
  irp_jpeg_out enc{};
  irp_result result{};
  irp_out_desc out{};
  int rc = 0;
  std::string error;
};

void Execute(napi_env, void* data) {
  Job* j = static_cast<Job*>(data);
  if (j->transcode) {   // capacity: half a byte per sample is ample at quality 85; one retry with the size the library reports
    int w = 0, h = 0, c = 0, ow = 0, oh = 0;
    j->rc = irp_jpeg_info(j->jpeg.data, j->jpeg.size, &w, &h, &c);
    if (j->rc == IRP_OK) j->rc = irp_preprocess_dims(w, h, j->jpeg.exif_orientation, &ow, &oh);
    for (int attempt = 0; j->rc == IRP_OK && attempt < 2; attempt++) {
      j->enc.capacity = attempt ? j->enc.size : static_cast<size_t>(ow) * oh * (c == 1 ? 1 : 3) / 2 + 8192;   // + header and a small profile
      std::free(j->enc.data);
      j->enc.data = static_cast<uint8_t*>(std::malloc(j->enc.capacity));
      if (!j->enc.data) { j->rc = IRP_ERR_NOMEM; break; }
      // queued like every other request: the dispatcher batches the uploads of concurrent HTTP requests
      irp_ticket ticket = nullptr;
      char err[256] = {0};
      j->rc = irp_submit_transcode(j->ctx, &j->jpeg, &j->result, j->quality, &j->enc, &ticket);
      if (j->rc == IRP_OK) j->rc = irp_wait(j->ctx, ticket, err, sizeof err);
      if (j->rc != IRP_OK) j->error = err[0] ? err : "irp request failed";
      if (j->rc != IRP_ERR_CAPACITY) break;
      if (attempt == 0) j->rc = IRP_OK;
    }
    if (j->rc == IRP_ERR_UNSUPPORTED) j->error = "unsupported";
    return;
  }
  if (j->preprocess) {
    int ow = 0, oh = 0;
    j->rc = irp_preprocess_dims(j->desc.width, j->desc.height, j->desc.exif_orientation, &ow, &oh);
    if (j->rc == IRP_OK) {
      const int oc = j->desc.channels == 1 ? 1 : 3;
      j->out.capacity = static_cast<size_t>(ow) * oh * oc;
      j->out.pixels = static_cast<uint8_t*>(std::malloc(j->out.capacity));
      j->out.pitch = 0;
      j->out.on_device = 0;
      if (!j->out.pixels) j->rc = IRP_ERR_NOMEM;
    }
  }
  if (j->rc == IRP_OK) {
    // one image per promise, several promises in flight (restorator.js:196-211): queue it and let the
    // context's dispatcher batch it with whatever the other libuv workers submitted meanwhile
    irp_ticket ticket = nullptr;
    char err[256] = {0};
    j->rc = j->file ? irp_submit_jpeg(j->ctx, &j->jpeg, &j->result, nullptr, &ticket)
                    : irp_submit(j->ctx, &j->desc, j->preprocess ? nullptr : &j->result, j->preprocess ? &j->out : nullptr, &ticket);
    if (j->rc == IRP_OK) j->rc = irp_wait(j->ctx, ticket, err, sizeof err);
    if (j->rc != IRP_OK) j->error = err[0] ? err : "irp request failed";
  } else {
    j->error = "unsupported geometry or out of memory";
  }
}

void Complete(napi_env env, napi_status, void* data) {
  Job* j = static_cast<Job*>(data);
  if (j->rc != IRP_OK) {  // whole-call failure == rejected promise, classifier.js:91-95
    napi_value msg, err;
    napi_create_string_utf8(env, j->error.c_str(), NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, nullptr, msg, &err);
    napi_reject_deferred(env, j->deferred, err);
    std::free(j->out.pixels);
    std::free(j->enc.data);
  } else if (j->transcode) {
    napi_value obj, buf, v, ab, arr;
    void *copy = nullptr, *dst = nullptr;
    napi_create_object(env, &obj);
    napi_create_buffer_copy(env, j->enc.size, j->enc.data, &copy, &buf);
    std::free(j->enc.data);
    napi_set_named_property(env, obj, "file", buf);
    napi_create_arraybuffer(env, sizeof(double) * IRP_NUM_SCORES, &dst, &ab);
    std::memcpy(dst, j->result.score, sizeof(double) * IRP_NUM_SCORES);
    napi_create_typedarray(env, napi_float64_array, IRP_NUM_SCORES, ab, 0, &arr);
    napi_set_named_property(env, obj, "scores", arr);
    napi_create_int32(env, j->enc.width, &v);
    napi_set_named_property(env, obj, "width", v);
    napi_create_int32(env, j->enc.height, &v);
    napi_set_named_property(env, obj, "height", v);
    napi_create_int32(env, j->enc.channels, &v);
    napi_set_named_property(env, obj, "channels", v);
    napi_resolve_deferred(env, j->deferred, obj);
  } else if (j->preprocess) {
    napi_value obj, buf, v;
    void* copy = nullptr;
    napi_create_object(env, &obj);
    napi_create_buffer_copy(env, j->out.capacity, j->out.pixels, &copy, &buf);
    std::free(j->out.pixels);
    napi_set_named_property(env, obj, "data", buf);
    napi_create_int32(env, j->out.width, &v);
    napi_set_named_property(env, obj, "width", v);
    napi_create_int32(env, j->out.height, &v);
    napi_set_named_property(env, obj, "height", v);
    napi_create_int32(env, j->out.channels, &v);
    napi_set_named_property(env, obj, "channels", v);
    napi_resolve_deferred(env, j->deferred, obj);
  } else {
    // Float64Array(11): the seven scores, then irp_result.issues (three IRP_ISSUE bytes or 255, and their count) —
    // PromptEnhancerService._identifyTopIssues' answer, already worked out by the library
    napi_value ab, arr;
    void* dst = nullptr;
    napi_create_arraybuffer(env, sizeof(double) * (IRP_NUM_SCORES + 4), &dst, &ab);
    std::memcpy(dst, j->result.score, sizeof(double) * IRP_NUM_SCORES);
    for (int k = 0; k < 4; k++) static_cast<double*>(dst)[IRP_NUM_SCORES + k] = j->result.issues[k];
    napi_create_typedarray(env, napi_float64_array, IRP_NUM_SCORES + 4, ab, 0, &arr);
    napi_resolve_deferred(env, j->deferred, arr);
  }
  napi_delete_reference(env, j->input_ref);
  napi_delete_async_work(env, j->work);
  delete j;
}

napi_value Submit(napi_env env, napi_callback_info info, bool preprocess) {
  size_t argc = 6;
  napi_value argv[6];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  Job* j = new Job();
  j->preprocess = preprocess;
  void* ctx = nullptr;
  napi_get_value_external(env, argv[0], &ctx);
  j->ctx = static_cast<irp_ctx*>(ctx);
  void* data = nullptr;
  size_t len = 0;
  napi_get_buffer_info(env, argv[1], &data, &len);
  napi_create_reference(env, argv[1], 1, &j->input_ref);  // JS owns the Buffer; keep it alive
  int32_t w = 0, h = 0, c = 0, last = 0;
  napi_get_value_int32(env, argv[2], &w);
  napi_get_value_int32(env, argv[3], &h);
  napi_get_value_int32(env, argv[4], &c);
  if (preprocess) {
    napi_get_value_int32(env, argv[5], &last);
  } else {
    bool b = false;
    napi_get_value_bool(env, argv[5], &b);
    last = b;
  }
  j->desc.pixels = static_cast<const uint8_t*>(data);
  j->desc.pitch = static_cast<size_t>(w) * c;
  j->desc.width = w;
  j->desc.height = h;
  j->desc.channels = c;
  j->desc.is_jpeg = preprocess ? 1 : last;
  j->desc.exif_orientation = preprocess ? last : 1;
  j->desc.on_device = 0;
  napi_value promise, name;
  napi_create_promise(env, &j->deferred, &promise);
  if (len < j->desc.pitch * static_cast<size_t>(h)) {
    napi_value msg, err;
    napi_create_string_utf8(env, "pixel buffer smaller than width*height*channels", NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, nullptr, msg, &err);
    napi_reject_deferred(env, j->deferred, err);
    napi_delete_reference(env, j->input_ref);
    delete j;
    return promise;
  }
  napi_create_string_utf8(env, preprocess ? "irp.preprocess" : "irp.analyze", NAPI_AUTO_LENGTH, &name);
  napi_create_async_work(env, nullptr, name, Execute, Complete, j, &j->work);
  napi_queue_async_work(env, j->work);
  return promise;
}

// analyzeFile(ctx, file:Buffer) -> Promise<Float64Array(7)>; rejects with "unsupported" for anything but a Huffman-coded 8-bit
// JPEG, in which case the shim decodes with sharp and calls analyzeRaw
napi_value AnalyzeFile(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  Job* j = new Job();
  j->file = true;
  void* ctx = nullptr;
  napi_get_value_external(env, argv[0], &ctx);
  j->ctx = static_cast<irp_ctx*>(ctx);
  void* data = nullptr;
  size_t len = 0;
  napi_get_buffer_info(env, argv[1], &data, &len);
  napi_create_reference(env, argv[1], 1, &j->input_ref);
  j->jpeg.data = static_cast<const uint8_t*>(data);
  j->jpeg.size = len;
  j->jpeg.exif_orientation = 1;   // the classifier never rotates (classifier.js has no .rotate())
  napi_value promise, name;
  napi_create_promise(env, &j->deferred, &promise);
  int w = 0, h = 0, c = 0;
  if (irp_jpeg_info(j->jpeg.data, len, &w, &h, &c) != IRP_OK) {
    napi_value msg, err;
    napi_create_string_utf8(env, "unsupported", NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, nullptr, msg, &err);
    napi_reject_deferred(env, j->deferred, err);
    napi_delete_reference(env, j->input_ref);
    delete j;
    return promise;
  }
  napi_create_string_utf8(env, "irp.analyzeFile", NAPI_AUTO_LENGTH, &name);
  napi_create_async_work(env, nullptr, name, Execute, Complete, j, &j->work);
  napi_queue_async_work(env, j->work);
  return promise;
}

// transcodeFile(ctx, file:Buffer, orientation, quality): rejects with "unsupported" for anything but a Huffman-coded 8-bit JPEG
napi_value TranscodeFile(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value argv[4];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  Job* j = new Job();
  j->transcode = true;
  void* ctx = nullptr;
  napi_get_value_external(env, argv[0], &ctx);
  j->ctx = static_cast<irp_ctx*>(ctx);
  void* data = nullptr;
  size_t len = 0;
  napi_get_buffer_info(env, argv[1], &data, &len);
  napi_create_reference(env, argv[1], 1, &j->input_ref);
  int32_t orientation = 1, quality = 85;
  if (argc > 2) napi_get_value_int32(env, argv[2], &orientation);
  if (argc > 3) napi_get_value_int32(env, argv[3], &quality);
  j->jpeg.data = static_cast<const uint8_t*>(data);
  j->jpeg.size = len;
  j->jpeg.exif_orientation = orientation;
  j->quality = quality;
  napi_value promise, name;
  napi_create_promise(env, &j->deferred, &promise);
  napi_create_string_utf8(env, "irp.transcodeFile", NAPI_AUTO_LENGTH, &name);
  napi_create_async_work(env, nullptr, name, Execute, Complete, j, &j->work);
  napi_queue_async_work(env, j->work);
  return promise;
}

napi_value SetOutputIcc(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  void* ctx = nullptr;
  napi_get_value_external(env, argv[0], &ctx);
  void* data = nullptr;
  size_t len = 0;
  if (argc > 1) napi_get_buffer_info(env, argv[1], &data, &len);   // not a Buffer (null / undefined): clears the profile
  if (irp_set_output_icc(static_cast<irp_ctx*>(ctx), static_cast<const uint8_t*>(data), data ? len : 0) != IRP_OK)
    napi_throw_error(env, nullptr, "bad ICC profile");
  return nullptr;
}

napi_value AnalyzeRaw(napi_env env, napi_callback_info info) { return Submit(env, info, false); }
napi_value PreprocessRaw(napi_env env, napi_callback_info info) { return Submit(env, info, true); }

void FinalizeCtx(napi_env, void* data, void*) { irp_destroy(static_cast<irp_ctx*>(data)); }

napi_value CreateContext(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  int32_t device = 0;
  if (argc > 0) napi_get_value_int32(env, argv[0], &device);
  irp_opts opts{};
  opts.struct_size = sizeof(opts);
  irp_ctx* ctx = irp_create(device, &opts);
  if (!ctx) {  // no CPU fallback: surface the error
    napi_throw_error(env, nullptr, irp_last_error(nullptr));
    return nullptr;
  }
  napi_value ext;
  napi_create_external(env, ctx, FinalizeCtx, nullptr, &ext);
  return ext;
}

napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor props[] = {
      {"createContext", nullptr, CreateContext, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"analyzeFile", nullptr, AnalyzeFile, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"transcodeFile", nullptr, TranscodeFile, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"setOutputIcc", nullptr, SetOutputIcc, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"analyzeRaw", nullptr, AnalyzeRaw, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"preprocessRaw", nullptr, PreprocessRaw, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  napi_define_properties(env, exports, 6, props);
  return exports;
}

}  // namespace

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
