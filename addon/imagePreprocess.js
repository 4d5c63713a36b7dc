// Drop-in for server-node/src/middleware/imagePreprocess.js — SOURCE ONLY / UNVERIFIED (no Node here).
// Same export, same req.file fields, same operation strings and the same problem documents; the sharp pipeline
// (.rotate() -> resize fit inside 2048 -> .jpeg(q85 4:4:4) -> ICC) is ONE native call when the upload is a JPEG the
// library takes (baseline or progressive): decoded, oriented, resized and re-encoded on the device (optimised Huffman tables as `mozjpeg: true`
// implies; no trellis quantisation, sequential instead of progressive scans).  Other containers are decoded by
// sharp once, processed natively as raw pixels and encoded by sharp as before.
import sharp from 'sharp';
import { createRequire } from 'node:module';
import { createProblem } from '../utils/problem.js';

const native = createRequire(import.meta.url)('./build/Release/irp_addon.node');
const MAX_DIMENSION = 2048;
const JPEG_QUALITY = 85;
const IRP_JPEG_OPTIMIZE = 0x100;
const IRP_JPEG_ICC_SRGB = 1 << 16;   // IRP_JPEG_ICC(IRP_ICC_SRGB): the library's generated sRGB profile, named per call

let ctx = null;
function context() {
  if (!ctx) {
    ctx = native.createContext(Number(process.env.IRP_DEVICE ?? 0));
  }
  return ctx;
}

const problem = (slug, title, status, detail) =>
  createProblem({ type: `https://docs.image-restoration.ai/problem/${slug}`, title, status, detail });

function targetSize(width, height) {
  if (Math.max(width, height) <= MAX_DIMENSION) return null;
  const scale = MAX_DIMENSION / Math.max(width, height);
  return { width: Math.round(width * scale), height: Math.round(height * scale) };
}

export async function preprocessImage(req, _res, next) {
  if (!req.file?.buffer) return next(problem('image-missing', 'Image File Required', 400, 'An image file must be provided in the request.'));
  try {
    const source = req.file.buffer;
    const meta = await sharp(source, { failOnError: false }).metadata();
    const operations = ['auto_orient'];
    const target = targetSize(meta.width, meta.height);
    if (target) operations.push(`resize_${target.width}x${target.height}`);
    let processed, outInfo;
    try {
      const r = await native.transcodeFile(context(), source, meta.orientation ?? 1, JPEG_QUALITY | IRP_JPEG_OPTIMIZE | IRP_JPEG_ICC_SRGB);   // 'attach_sRGB_icc' is always true
      processed = r.file;
      outInfo = { width: r.width, height: r.height, channels: r.channels };
    } catch (e) {
      if (e.message !== 'unsupported') throw e;
      const { data, info } = await sharp(source, { failOnError: false }).raw().toBuffer({ resolveWithObject: true });
      const out = await native.preprocessRaw(context(), data, info.width, info.height, info.channels, meta.orientation ?? 1);
      processed = await sharp(out.data, { raw: { width: out.width, height: out.height, channels: out.channels } })
        .jpeg({ quality: JPEG_QUALITY, chromaSubsampling: '4:4:4', mozjpeg: true }).withMetadata({ icc: 'sRGB' }).toBuffer();
      outInfo = { width: out.width, height: out.height, channels: out.channels };
    }
    operations.push(`compress_jpeg_q${JPEG_QUALITY}`, 'attach_sRGB_icc');
    Object.assign(req.file, {   // the fields the reference middleware rewrites (imagePreprocess.js:70-78)
      originalBuffer: source, originalMetadata: meta, buffer: processed, size: processed.length,
      processedMetadata: { ...outInfo, format: 'jpeg' }, preprocessOperations: operations,
      mimetype: 'image/jpeg', detectedMime: 'image/jpeg', detectedExt: 'jpg',
    });
    return next();
  } catch (error) {
    return next(problem('preprocess-failed', 'Image Preprocessing Failed', 422, error.message || 'Unable to preprocess the uploaded image.'));
  }
}
