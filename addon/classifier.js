// Drop-in for server-node/src/services/classifier.js — SOURCE ONLY / UNVERIFIED (no Node here).
// Same exports, same analyze() key order, same logging and span calls; the six sharp pipelines
// and the JS reductions are replaced by ONE native call.  JPEG files (baseline or progressive) are handed over as they
// are (decoded on the device); other containers are decoded by sharp once instead of six times.
import { trace, SpanStatusCode } from '@opentelemetry/api';
import sharp from 'sharp';
import { createRequire } from 'node:module';

const native = createRequire(import.meta.url)('./build/Release/irp_addon.node');
const KEYS = ['blur', 'noise', 'lowLight', 'compression', 'scratch', 'fade', 'colorShift'];

const DEGRADATION_TYPES = {
  blur: 'Motion blur or out-of-focus areas',
  noise: 'Grain and digital noise',
  lowLight: 'Underexposed or shadow detail loss',
  compression: 'JPEG artifacts and quality loss',
  scratch: 'Physical damage and blemishes',
  fade: 'Color loss and contrast reduction',
  colorShift: 'White balance and color cast issues',
};

let ctx = null;

export class ClassifierService {
  constructor({ logger, device = Number(process.env.IRP_DEVICE ?? 0) } = {}) {
    this.logger = logger ?? console;
    this.device = device;
  }

  async analyze(imageBuffer) {
    const span = trace.getTracer('classifier').startSpan('classifier.analyze', {
      attributes: { 'image.size_bytes': imageBuffer.length, 'classifier.version': '1.0.0-b200' },
    });
    try {
      ctx ??= native.createContext(this.device);
      const metadata = await sharp(imageBuffer).metadata();
      let scores;
      try {
        // a JPEG (baseline or progressive) goes to the GPU as it is: decoded there bit-exactly as libjpeg-turbo (sharp's decoder) would
        scores = await native.analyzeFile(ctx, imageBuffer);
      } catch (e) {
        if (e.message !== 'unsupported') throw e;
        // PNG / WebP (and JPEG kinds the device refuses): one sharp decode (instead of six), then the raw entry point
        const { data, info } = await sharp(imageBuffer).raw().toBuffer({ resolveWithObject: true });
        scores = await native.analyzeRaw(ctx, data, info.width, info.height, info.channels, metadata.format === 'jpeg');
      }
      const analysis = Object.fromEntries(KEYS.map((k, i) => [k, scores[i]]));
      // scores[7..10] = irp_result.issues: the three highest scores above 0.3 with their severity, already sorted
      // (promptEnhancer.js:121-145); kept on the instance for PromptEnhancerService, the returned object stays the 7 keys
      const SEVERITY = [null, 'low', 'medium', 'high'];
      const issues = [];
      for (let k = 0; k < scores[10]; k++) {
        const v = scores[7 + k];
        issues.push({ type: KEYS[v & 15], confidence: scores[v & 15], severity: SEVERITY[v >> 4] });
      }
      this.lastTopIssues = issues;
      const topIssues = issues.map((i) => [i.type, i.confidence]);
      span.setAttributes({
        'image.width': metadata.width, 'image.height': metadata.height, 'image.format': metadata.format,
        'image.channels': metadata.channels,
        'classifier.top_issues': topIssues.map(([t, s]) => `${t}:${s.toFixed(2)}`).join(','),
        'classifier.issue_count': topIssues.length,
      });
      this.logger.debug('[classifier] Analysis complete', {
        topIssues: topIssues.map(([type, score]) => ({ type, score: score.toFixed(2) })),
        imageSize: `${metadata.width}x${metadata.height}`,
      });
      span.setStatus({ code: SpanStatusCode.OK });
      return analysis;
    } catch (error) {
      span.recordException(error);
      span.setStatus({ code: SpanStatusCode.ERROR, message: error.message });
      this.logger.error('[classifier] Analysis failed', { error: error.message });
      throw error;
    } finally {
      span.end();
    }
  }

  static getDegradationTypes() {
    return { ...DEGRADATION_TYPES };
  }
}

export function createClassifierService(options = {}) {
  return new ClassifierService(options);
}
