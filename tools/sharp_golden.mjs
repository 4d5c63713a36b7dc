#!/usr/bin/env node
// sharp_golden.mjs — pin the CPU oracle against the REAL reference arithmetic (sharp 0.33.5 / libvips 8.15).
//
// This image has no Node, so the repository's parity is "CUDA == oracle" with the oracle's libvips half restated
// from memory (oracle/irp_oracle.c header).  A maintainer with Node closes that gap in one command:
//
//   python tools/sharp_fixtures.py /tmp/irp_fixtures
//   node tools/sharp_golden.mjs /path/to/image-restoration-platform /tmp/irp_fixtures > tests/golden/sharp_golden.json
//   python -m pytest tests/test_sharp_golden.py -q        # says which switch (luma / coef / blur / reduce mode) fits
//
// Everything below is the reference's own code or its literal calls:
//   * ClassifierService.analyze and _detectBlockiness are IMPORTED from the reference (services/classifier.js:40-99,
//     288-308) and run on each fixture;
//   * the intermediate buffers come from the same sharp pipelines that file builds, call for call
//     (classifier.js:51-52 metadata / stats, :107-115 Lap8, :135-143 Sharp9, :199-207 Lap4, :296-297 raw and blur(1));
//   * preprocessImage is IMPORTED from the reference (middleware/imagePreprocess.js:24-91) and run as Express would;
//     the same pipeline is run once more ending in .raw() instead of .jpeg(), because the resize arithmetic is judged on
//     pixels (SURVEY.md section 8a row P4: "the GPU stage ends at normalised raw u8 RGB").
// Buffers travel as base64 (cases are sized for that: tests/golden/sharp_cases.py).
import { createRequire } from 'node:module';
import { createHash } from 'node:crypto';
import { readFileSync } from 'node:fs';
import path from 'node:path';
import { pathToFileURL } from 'node:url';

const [refRoot, fixtureDir] = process.argv.slice(2);
if (!refRoot || !fixtureDir) {
  console.error('usage: node tools/sharp_golden.mjs <reference repo root> <fixture dir>  > tests/golden/sharp_golden.json');
  process.exit(2);
}
const server = path.join(refRoot, 'server-node');
const require = createRequire(path.join(server, 'package.json'));
const sharp = require('sharp');   // the reference's own pinned copy (server-node/package.json:37)
const { ClassifierService } = await import(pathToFileURL(path.join(server, 'src/services/classifier.js')).href);
const { preprocessImage } = await import(pathToFileURL(path.join(server, 'src/middleware/imagePreprocess.js')).href);

const quiet = { debug() {}, info() {}, warn() {}, error() {} };
const sha = (b) => createHash('sha256').update(b).digest('hex');
const b64 = (b) => Buffer.from(b).toString('base64');
// exact integer moments of a byte buffer (BigInt: a 12 MP buffer's sum of squares passes 2^53)
function moments(buf) {
  let s = 0n, q = 0n;
  for (const v of buf) { s += BigInt(v); q += BigInt(v * v); }
  return { n: buf.length, sum: s.toString(), sumsq: q.toString(), sha256: sha(buf) };
}
const pack = (buf, keep) => ({ ...moments(buf), ...(keep ? { base64: b64(buf) } : {}) });

const manifest = JSON.parse(readFileSync(path.join(fixtureDir, 'manifest.json'), 'utf8'));
const out = {
  versions: { sharp: require('sharp/package.json').version, ...sharp.versions, simd: sharp.simd?.() ?? null, node: process.version },
  cases: [],
};
const service = new ClassifierService({ logger: quiet });

for (const m of manifest) {
  const file = readFileSync(path.join(fixtureDir, m.file));
  const rec = { name: m.name, group: m.group, width: m.width, height: m.height, channels: m.channels, orientation: m.orientation,
                pixels_sha256: m.pixels_sha256 };
  const keep = m.width * m.height <= 70000;   // whole buffers for the small cases
  // the decoded pixels sharp works from must be the seeded ones
  const raw = await sharp(file).raw().toBuffer({ resolveWithObject: true });   // classifier.js:296
  rec.decoded = { info: raw.info, sha256: sha(raw.data) };

  if (m.group === 'classify') {
    const metadata = await sharp(file).metadata();                             // classifier.js:51
    const stats = await sharp(file).stats();                                   // classifier.js:52
    rec.metadata = { width: metadata.width, height: metadata.height, channels: metadata.channels, format: metadata.format };
    rec.stats = stats.channels.map((c) => ({ min: c.min, max: c.max, sum: c.sum, squaresSum: c.squaresSum, mean: c.mean, stdev: c.stdev }));
    rec.grey = pack(await sharp(file).grayscale().raw().toBuffer(), keep);      // G1 alone
    const conv = (kernel) => sharp(file).grayscale().convolve({ width: 3, height: 3, kernel }).raw().toBuffer();
    rec.lap8 = pack(await conv([-1, -1, -1, -1, 8, -1, -1, -1, -1]), keep);     // classifier.js:107-115
    rec.sharp9 = pack(await conv([-1, -1, -1, -1, 9, -1, -1, -1, -1]), keep);   // classifier.js:135-143
    const lap4 = await conv([0, -1, 0, -1, 4, -1, 0, -1, 0]);                   // classifier.js:199-207
    rec.lap4 = pack(lap4, keep);
    rec.scratch_indicator = service._detectLinearFeatures(lap4, metadata.width, metadata.height);   // classifier.js:310-337
    rec.original = pack(raw.data, false);                                       // classifier.js:296
    rec.blurred = pack((await sharp(file).blur(1).raw().toBuffer({ resolveWithObject: true })).data, keep);   // classifier.js:297
    rec.analysis = await service.analyze(file);                                 // the seven scores (format png: compression 0)
    rec.blockiness = await service._detectBlockiness(file);                     // what a JPEG of these pixels would score
  } else {
    // the middleware as Express runs it (imagePreprocess.js:24-91)
    const req = { file: { buffer: file, mimetype: 'image/png', originalname: m.file, size: file.length } };
    const err = await new Promise((resolve) => preprocessImage(req, {}, resolve));
    if (err) {
      rec.error = { status: err.status, title: err.title, detail: err.detail ?? String(err) };
    } else {
      rec.operations = req.file.preprocessOperations;
      rec.processedMetadata = { width: req.file.processedMetadata.width, height: req.file.processedMetadata.height,
                                channels: req.file.processedMetadata.channels, format: req.file.processedMetadata.format,
                                isProgressive: req.file.processedMetadata.isProgressive, chromaSubsampling: req.file.processedMetadata.chromaSubsampling,
                                hasProfile: req.file.processedMetadata.hasProfile };
      rec.file = { size: req.file.buffer.length, sha256: sha(req.file.buffer), base64: b64(req.file.buffer) };
      // what a decoder makes of that file (for the decoded-pixel comparison of the encoder row)
      const dec = await sharp(req.file.buffer).raw().toBuffer({ resolveWithObject: true });
      rec.file_decoded = { info: dec.info, ...pack(dec.data, true) };
    }
    // the same pipeline up to the encoder: .rotate() + .resize() of imagePreprocess.js:42-53, pixels out
    const meta = await sharp(file, { failOnError: false }).metadata();
    let p = sharp(file, { failOnError: false }).rotate();
    const MAX = 2048;
    if (meta.width > MAX || meta.height > MAX) {
      const scale = MAX / Math.max(meta.width, meta.height);
      p = p.resize({ width: Math.round(meta.width * scale), height: Math.round(meta.height * scale), fit: 'inside', withoutEnlargement: true });
    }
    const px = await p.raw().toBuffer({ resolveWithObject: true });
    rec.resized = { info: px.info, ...pack(px.data, true) };
  }
  out.cases.push(rec);
}
process.stdout.write(JSON.stringify(out));
