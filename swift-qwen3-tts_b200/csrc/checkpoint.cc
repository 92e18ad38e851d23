// See checkpoint.hpp.  Reference rules restated (not copied) from
// Sources/Qwen3TTS/Models/Qwen3.swift:1246-1260, 1498-1512, 1530-1543, 1581-1588, 1687-1724
// and Sources/Qwen3TTS/Models/Config.swift:361-409, 574-592.
#include "checkpoint.hpp"

#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

#include "json_min.hpp"

namespace q3 {

bool is_mlx_conv_layout(const std::vector<int64_t>& s) {
  if (s.size() != 3) return false;
  int64_t d2 = s[1], d3 = s[2];
  if (d2 == 1) return d3 > 64;
  if (d3 == 1) return d2 <= 64;
  return d2 < d3;
}

static std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(Q3TTS_EIO, "cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

static int geti(const Json& o, const char* k, int dflt) {
  const Json* v = o.find(k);
  if (!v || v->is_null()) return dflt;
  if (v->kind != Json::Num) throw Error(Q3TTS_EFORMAT, std::string("config key '") + k + "' is not a number");
  return (int)v->num;
}
static float getf(const Json& o, const char* k, float dflt) {
  const Json* v = o.find(k);
  if (!v || v->is_null()) return dflt;
  if (v->kind != Json::Num) throw Error(Q3TTS_EFORMAT, std::string("config key '") + k + "' is not a number");
  return (float)v->num;
}
static void get_list(const Json& o, const char* k, const std::vector<int>& dflt, int32_t* n, int32_t* out) {
  std::vector<int> v = dflt;
  const Json* j = o.find(k);
  if (j && !j->is_null()) {
    if (j->kind != Json::Arr) throw Error(Q3TTS_EFORMAT, std::string("config key '") + k + "' is not a list");
    v.clear();
    for (auto& e : j->arr) {
      if (e.kind != Json::Num) throw Error(Q3TTS_EFORMAT, std::string("config key '") + k + "' has a non-number");
      v.push_back((int)e.num);
    }
  }
  if (v.size() > 8) throw Error(Q3TTS_EFORMAT, std::string("config key '") + k + "' has more than 8 entries");
  *n = (int32_t)v.size();
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}

void parse_tokenizer_config(const std::string& dir, q3tts_config* c) {
  std::string text = read_file(dir + "/config.json");
  Json root;
  try {
    root = JsonParser(text.data(), text.size()).parse();
  } catch (const std::exception& e) {
    throw Error(Q3TTS_EFORMAT, std::string("config.json: ") + e.what());
  }
  if (root.kind != Json::Obj) throw Error(Q3TTS_EFORMAT, "config.json: top level is not an object");
  std::memset(c, 0, sizeof(*c));
  c->decode_upsample_rate = geti(root, "decode_upsample_rate", 1920);   // Cfg.swift:590
  c->output_sample_rate = geti(root, "output_sample_rate", 24000);      // Cfg.swift:589
  const Json* enc = root.find("encoder_config");
  c->has_encoder_config = (enc && !enc->is_null()) ? 1 : 0;             // ST.swift:808-817
  const Json* dc = root.find("decoder_config");
  if (!dc || dc->kind != Json::Obj)                                     // ST.swift:801-805 (fatalError there)
    throw Error(Q3TTS_EFORMAT, "config.json: decoder_config is required");
  const Json& d = *dc;                                                  // Cfg.swift:388-408
  c->latent_dim = geti(d, "latent_dim", 1024);
  c->codebook_dim = geti(d, "codebook_dim", 512);
  c->codebook_size = geti(d, "codebook_size", 2048);
  c->decoder_dim = geti(d, "decoder_dim", 1536);
  c->hidden_size = geti(d, "hidden_size", 512);
  c->intermediate_size = geti(d, "intermediate_size", 1024);
  c->num_hidden_layers = geti(d, "num_hidden_layers", 8);
  c->num_attention_heads = geti(d, "num_attention_heads", 16);
  c->num_key_value_heads = geti(d, "num_key_value_heads", 16);
  c->head_dim = geti(d, "head_dim", 64);
  c->rms_norm_eps = getf(d, "rms_norm_eps", 1e-5f);
  c->rope_theta = getf(d, "rope_theta", 10000.0f);
  c->sliding_window = geti(d, "sliding_window", 72);
  c->num_quantizers = geti(d, "num_quantizers", 16);
  c->num_semantic_quantizers = geti(d, "num_semantic_quantizers", 1);
  c->semantic_codebook_size = geti(d, "semantic_codebook_size", 4096);
  get_list(d, "upsample_rates", {8, 5, 4, 3}, &c->num_upsample_rates, c->upsample_rates);
  get_list(d, "upsampling_ratios", {2, 2}, &c->num_upsampling_ratios, c->upsampling_ratios);
  c->layer_scale_initial_scale = getf(d, "layer_scale_initial_scale", 0.01f);
  int tot = 1;
  for (int i = 0; i < c->num_upsample_rates; ++i) tot *= c->upsample_rates[i];
  for (int i = 0; i < c->num_upsampling_ratios; ++i) tot *= c->upsampling_ratios[i];
  c->total_upsample = tot;                                              // Cfg.swift:411-414
  // structural checks the reference leaves implicit
  if (c->num_upsample_rates != 4)
    throw Error(Q3TTS_EFORMAT, "decoder_config.upsample_rates must have 4 entries (MainDecoder has block0..3, ST.swift:654-670)");
  if (c->num_semantic_quantizers < 1 || c->num_quantizers <= c->num_semantic_quantizers || c->num_quantizers > 64)
    throw Error(Q3TTS_EFORMAT, "decoder_config: bad num_quantizers / num_semantic_quantizers");
  if (c->codebook_dim % 2 || c->decoder_dim % 16 || c->num_attention_heads * c->head_dim <= 0 ||
      c->num_key_value_heads <= 0 || c->num_attention_heads % c->num_key_value_heads)
    throw Error(Q3TTS_EFORMAT, "decoder_config: inconsistent dimensions");
}

std::map<std::string, std::vector<int64_t>> expected_decoder_tensors(const q3tts_config& c) {
  std::map<std::string, std::vector<int64_t>> m;
  const int64_t half = c.codebook_dim / 2, L = c.latent_dim, H = c.hidden_size, I = c.intermediate_size;
  const int64_t A = (int64_t)c.num_attention_heads * c.head_dim, KV = (int64_t)c.num_key_value_heads * c.head_dim;
  const std::string q = "decoder.quantizer.";
  for (int i = 0; i < c.num_semantic_quantizers; ++i)
    m[q + "rvq_first.vq.layers." + std::to_string(i) + ".codebook.embed.weight"] = {c.semantic_codebook_size, half};
  for (int i = 0; i < c.num_quantizers - c.num_semantic_quantizers; ++i)
    m[q + "rvq_rest.vq.layers." + std::to_string(i) + ".codebook.embed.weight"] = {c.codebook_size, half};
  for (const char* part : {"rvq_first", "rvq_rest"})
    m[q + part + ".output_proj.weight"] = {c.codebook_dim, 1, half};
  m["decoder.pre_conv.conv.weight"] = {L, 3, c.codebook_dim};
  m["decoder.pre_conv.conv.bias"] = {L};
  const std::string pt = "decoder.pre_transformer.";
  m[pt + "input_proj.weight"] = {H, L};
  m[pt + "input_proj.bias"] = {H};
  m[pt + "output_proj.weight"] = {L, H};
  m[pt + "output_proj.bias"] = {L};
  m[pt + "norm.weight"] = {H};
  for (int n = 0; n < c.num_hidden_layers; ++n) {
    std::string p = pt + "layers." + std::to_string(n) + ".";
    m[p + "self_attn.q_proj.weight"] = {A, H};
    m[p + "self_attn.k_proj.weight"] = {KV, H};
    m[p + "self_attn.v_proj.weight"] = {KV, H};
    m[p + "self_attn.o_proj.weight"] = {H, A};
    m[p + "mlp.gate_proj.weight"] = {I, H};
    m[p + "mlp.up_proj.weight"] = {I, H};
    m[p + "mlp.down_proj.weight"] = {H, I};
    m[p + "input_layernorm.weight"] = {H};
    m[p + "post_attention_layernorm.weight"] = {H};
    m[p + "self_attn_layer_scale.scale"] = {H};
    m[p + "mlp_layer_scale.scale"] = {H};
  }
  for (int i = 0; i < c.num_upsampling_ratios; ++i) {
    std::string u = "decoder.upsample." + std::to_string(i) + ".";
    int64_t r = c.upsampling_ratios[i];
    m[u + "0.conv.weight"] = {L, r, L};
    m[u + "0.conv.bias"] = {L};
    m[u + "1.dwconv.conv.weight"] = {L, 7, 1};
    m[u + "1.dwconv.conv.bias"] = {L};
    m[u + "1.norm.weight"] = {L};
    m[u + "1.norm.bias"] = {L};
    m[u + "1.pwconv1.weight"] = {4 * L, L};
    m[u + "1.pwconv1.bias"] = {4 * L};
    m[u + "1.pwconv2.weight"] = {L, 4 * L};
    m[u + "1.pwconv2.bias"] = {L};
    m[u + "1.gamma"] = {L};
  }
  const std::string dd = "decoder.decoder.";
  const int64_t D = c.decoder_dim;
  m[dd + "initConv.conv.weight"] = {D, 7, L};
  m[dd + "initConv.conv.bias"] = {D};
  for (int i = 0; i < c.num_upsample_rates; ++i) {
    int64_t cin = D >> i, cout = D >> (i + 1), r = c.upsample_rates[i];
    std::string b = dd + "block" + std::to_string(i) + ".";
    m[b + "snake.alpha"] = {cin};
    m[b + "snake.beta"] = {cin};
    m[b + "upsample.conv.weight"] = {cout, 2 * r, cin};
    m[b + "upsample.conv.bias"] = {cout};
    for (const char* res : {"res1", "res2", "res3"}) {
      std::string p = b + res + ".";
      m[p + "act1.alpha"] = {cout};
      m[p + "act1.beta"] = {cout};
      m[p + "conv1.conv.weight"] = {cout, 7, cout};
      m[p + "conv1.conv.bias"] = {cout};
      m[p + "act2.alpha"] = {cout};
      m[p + "act2.beta"] = {cout};
      m[p + "conv2.conv.weight"] = {cout, 1, cout};
      m[p + "conv2.conv.bias"] = {cout};
    }
  }
  int64_t cl = D >> c.num_upsample_rates;
  m[dd + "outSnake.alpha"] = {cl};
  m[dd + "outSnake.beta"] = {cl};
  m[dd + "outConv.conv.weight"] = {1, 7, cl};
  m[dd + "outConv.conv.bias"] = {1};
  return m;
}

// ---- safetensors -----------------------------------------------------------------------------
namespace {

struct Mapped {
  int fd = -1;
  const uint8_t* p = nullptr;
  size_t n = 0;
  ~Mapped() {
    if (p) munmap((void*)p, n);
    if (fd >= 0) close(fd);
  }
};

float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400u));
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
    }
  } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
  else bits = sign | ((exp + 112u) << 23) | (man << 13);
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}

void to_float(const uint8_t* src, const std::string& dtype, int64_t n, std::vector<float>* dst) {
  dst->resize((size_t)n);
  if (dtype == "F32") std::memcpy(dst->data(), src, (size_t)n * 4);
  else if (dtype == "F16") {
    const uint16_t* s = (const uint16_t*)src;
    for (int64_t i = 0; i < n; ++i) (*dst)[i] = half_to_float(s[i]);
  } else if (dtype == "BF16") {
    const uint16_t* s = (const uint16_t*)src;
    for (int64_t i = 0; i < n; ++i) {
      uint32_t b = (uint32_t)s[i] << 16;
      std::memcpy(&(*dst)[i], &b, 4);
    }
  } else if (dtype == "F64") {
    const double* s = (const double*)src;
    for (int64_t i = 0; i < n; ++i) (*dst)[i] = (float)s[i];
  } else throw Error(Q3TTS_EFORMAT, "unsupported safetensors dtype " + dtype);
}

size_t dtype_size(const std::string& d) {
  if (d == "F32") return 4;
  if (d == "F16" || d == "BF16") return 2;
  if (d == "F64") return 8;
  return 0;
}

HostTensor permute3(const HostTensor& t, int a0, int a1, int a2) {
  const int64_t s[3] = {t.shape[0], t.shape[1], t.shape[2]};
  const int ax[3] = {a0, a1, a2};
  HostTensor o;
  o.shape = {s[a0], s[a1], s[a2]};
  o.data.resize(t.data.size());
  const int64_t st[3] = {s[1] * s[2], s[2], 1};
  int64_t idx = 0;
  for (int64_t i = 0; i < o.shape[0]; ++i)
    for (int64_t j = 0; j < o.shape[1]; ++j)
      for (int64_t k = 0; k < o.shape[2]; ++k) o.data[idx++] = t.data[i * st[ax[0]] + j * st[ax[1]] + k * st[ax[2]]];
  return o;
}

bool has(const std::string& s, const char* sub) { return s.find(sub) != std::string::npos; }
bool starts(const std::string& s, const char* pre) { return s.rfind(pre, 0) == 0; }
void replace_all(std::string& s, const std::string& from, const std::string& to) {
  size_t pos = 0;
  while ((pos = s.find(from, pos)) != std::string::npos) {
    s.replace(pos, from.size(), to);
    pos += to.size();
  }
}

// `needle`: keep keys that START with it (contains == false) or CONTAIN it; `dtypes` (optional) receives the on-disk dtype names.
void read_safetensors(const std::string& path, TensorMap* raw, const char* needle = "decoder.", bool contains = false,
                      std::map<std::string, std::string>* dtypes = nullptr) {
  Mapped m;
  m.fd = open(path.c_str(), O_RDONLY);
  if (m.fd < 0) throw Error(Q3TTS_EIO, "cannot open " + path);
  struct stat st;
  if (fstat(m.fd, &st) != 0 || st.st_size < 8) throw Error(Q3TTS_EFORMAT, path + ": too small for safetensors");
  m.n = (size_t)st.st_size;
  void* p = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, m.fd, 0);
  if (p == MAP_FAILED) throw Error(Q3TTS_EIO, "mmap failed for " + path);
  m.p = (const uint8_t*)p;
  uint64_t hlen;
  std::memcpy(&hlen, m.p, 8);
  if (hlen > m.n - 8 || hlen > (1ull << 30)) throw Error(Q3TTS_EFORMAT, path + ": bad safetensors header length");
  Json hdr;
  try {
    hdr = JsonParser((const char*)m.p + 8, (size_t)hlen).parse();
  } catch (const std::exception& e) {
    throw Error(Q3TTS_EFORMAT, path + ": " + e.what());
  }
  if (hdr.kind != Json::Obj) throw Error(Q3TTS_EFORMAT, path + ": header is not an object");
  const uint8_t* base = m.p + 8 + hlen;
  const size_t payload = m.n - 8 - (size_t)hlen;
  for (auto& kv : hdr.obj) {
    const std::string& key = kv.first;
    if (key == "__metadata__") continue;
    if (contains ? !has(key, needle) : !starts(key, needle)) continue;  // decoder load: encoder.* is out of scope (and absent from the lite variant)
    const Json& e = kv.second;
    const Json *jd = e.find("dtype"), *js = e.find("shape"), *jo = e.find("data_offsets");
    if (!jd || !js || !jo || jd->kind != Json::Str || js->kind != Json::Arr || jo->kind != Json::Arr || jo->arr.size() != 2)
      throw Error(Q3TTS_EFORMAT, path + ": malformed entry for " + key);
    HostTensor t;
    for (auto& d : js->arr) t.shape.push_back((int64_t)d.num);
    const int64_t n = t.numel();
    const size_t b = (size_t)jo->arr[0].num, en = (size_t)jo->arr[1].num, es = dtype_size(jd->str);
    if (es == 0) throw Error(Q3TTS_EFORMAT, path + ": unsupported dtype " + jd->str + " for " + key);
    if (en < b || en > payload || (en - b) != (size_t)n * es)
      throw Error(Q3TTS_EFORMAT, path + ": data_offsets out of range for " + key);
    to_float(base + b, jd->str, n, &t.data);
    if (dtypes) (*dtypes)[key] = jd->str;
    (*raw)[key] = std::move(t);
  }
}

}  // namespace

static std::vector<std::string> list_safetensors(const std::string& dir) {
  std::vector<std::string> files;
  DIR* d = opendir(dir.c_str());
  if (!d) throw Error(Q3TTS_EIO, "cannot list " + dir);
  while (dirent* e = readdir(d)) {
    std::string n = e->d_name;
    if (n.size() > 12 && n.substr(n.size() - 12) == ".safetensors") files.push_back(dir + "/" + n);
  }
  closedir(d);
  std::sort(files.begin(), files.end());
  if (files.empty()) throw Error(Q3TTS_EIO, "no *.safetensors in " + dir);
  return files;
}

void load_codec_embeddings(const std::string& model_dir, CodecEmbeddingTables* out) {
  TensorMap raw;
  std::map<std::string, std::string> dtypes;
  for (auto& f : list_safetensors(model_dir)) read_safetensors(f, &raw, "codec_embedding", true, &dtypes);
  const std::string k0 = "talker.model.codec_embedding.weight";
  auto it = raw.find(k0);
  if (it == raw.end()) throw Error(Q3TTS_EFORMAT, model_dir + ": no " + k0);
  if (it->second.shape.size() != 2) throw Error(Q3TTS_EFORMAT, k0 + " must be 2-D");
  out->hidden = it->second.shape[1];
  out->dtype = dtypes[k0];
  out->tables.clear();
  out->tables.push_back(std::move(it->second));
  for (int i = 0;; ++i) {
    const std::string k = "talker.code_predictor.model.codec_embedding." + std::to_string(i) + ".weight";
    auto jt = raw.find(k);
    if (jt == raw.end()) break;
    if (jt->second.shape.size() != 2 || jt->second.shape[1] != out->hidden)
      throw Error(Q3TTS_EFORMAT, k + ": expected [vocab, " + std::to_string(out->hidden) + "]");
    if (dtypes[k] != out->dtype) throw Error(Q3TTS_EFORMAT, k + ": dtype differs from " + k0);
    out->tables.push_back(std::move(jt->second));
  }
  if (out->tables.size() < 2) throw Error(Q3TTS_EFORMAT, model_dir + ": no talker.code_predictor.model.codec_embedding.N.weight");
}

void load_checkpoint(const std::string& dir, Checkpoint* out) {
  parse_tokenizer_config(dir, &out->cfg);
  std::vector<std::string> files;
  DIR* d = opendir(dir.c_str());
  if (!d) throw Error(Q3TTS_EIO, "cannot list " + dir);
  while (dirent* e = readdir(d)) {
    std::string n = e->d_name;
    if (n.size() > 12 && n.substr(n.size() - 12) == ".safetensors") files.push_back(dir + "/" + n);
  }
  closedir(d);
  std::sort(files.begin(), files.end());
  if (files.empty()) throw Error(Q3TTS_EIO, "no *.safetensors in " + dir);
  TensorMap raw;
  for (auto& f : files) read_safetensors(f, &raw);  // later files win, like merge {_, new in new} (Q3.swift:1478)
  out->cfg.num_decoder_tensors = (int64_t)raw.size();

  static const std::pair<const char*, const char*> kIndexMap[] = {
      {"decoder.decoder.0", "decoder.decoder.initConv"}, {"decoder.decoder.1", "decoder.decoder.block0"},
      {"decoder.decoder.2", "decoder.decoder.block1"},   {"decoder.decoder.3", "decoder.decoder.block2"},
      {"decoder.decoder.4", "decoder.decoder.block3"},   {"decoder.decoder.5", "decoder.decoder.outSnake"},
      {"decoder.decoder.6", "decoder.decoder.outConv"}};
  TensorMap& san = out->tensors;
  std::map<std::string, std::pair<const HostTensor*, const HostTensor*>> books;  // base -> (usage, sum)
  for (auto& kv : raw) {
    const std::string& key = kv.first;
    const HostTensor& value = kv.second;
    if (has(key, "._codebook.cluster_usage") || has(key, "._codebook.embedding_sum")) {
      std::string base = key.substr(0, key.find("._codebook."));
      if (has(key, "cluster_usage")) books[base].first = &value;
      else books[base].second = &value;
      continue;
    }
    std::string nk = key;
    for (auto& im : kIndexMap)
      if (starts(key, im.first)) {
        nk = key;
        replace_all(nk, im.first, im.second);
        break;
      }
    replace_all(nk, ".block.0.", ".snake.");
    replace_all(nk, ".block.1.", ".upsample.");
    replace_all(nk, ".block.2.", ".res1.");
    replace_all(nk, ".block.3.", ".res2.");
    replace_all(nk, ".block.4.", ".res3.");
    HostTensor nv;
    bool set = false;
    const bool is3 = value.shape.size() == 3;
    const bool is_proj = (has(nk, "input_proj.weight") || has(nk, "output_proj.weight")) && has(nk, "quantizer");
    if (is_proj && is3) {  // unconditional [o,i,1] -> [o,1,i]
      nv = permute3(value, 0, 2, 1);
      set = true;
    }
    if (has(nk, "conv.weight") && is3 && !is_proj && !is_mlx_conv_layout(value.shape)) {
      nv = permute3(value, 0, 2, 1);
      set = true;
    }
    const bool is_tconv = (has(nk, "upsample") && has(nk, ".0.conv.weight")) ||
                          (has(nk, "decoder.decoder.block") && has(nk, "upsample.conv.weight"));
    if (is_tconv && is3) {  // overrides the generic branch, from the ORIGINAL value (Q3.swift:1704-1711)
      if (!is_mlx_conv_layout(value.shape)) nv = permute3(value, 1, 2, 0);
      else nv = value;
      set = true;
    }
    san[nk] = set ? std::move(nv) : value;
  }
  for (auto& b : books) {  // Q3.swift:1716-1724
    if (!b.second.first || !b.second.second) continue;
    const HostTensor& usage = *b.second.first;
    const HostTensor& sum = *b.second.second;
    if (usage.shape.size() != 1 || sum.shape.size() != 2 || sum.shape[0] != usage.shape[0])
      throw Error(Q3TTS_EFORMAT, "codebook tensors of " + b.first + " have inconsistent shapes");
    HostTensor e;
    e.shape = sum.shape;
    e.data.resize(sum.data.size());
    const int64_t D = sum.shape[1];
    for (int64_t r = 0; r < sum.shape[0]; ++r) {
      float den = std::min(std::max(usage.data[r], 1e-5f), FLT_MAX);
      for (int64_t c = 0; c < D; ++c) e.data[r * D + c] = sum.data[r * D + c] / den;
    }
    san[b.first + ".codebook.embed.weight"] = std::move(e);
  }
  // strict validation (the reference's verify: [] accepts anything, Q3.swift:1486)
  int64_t params = 0;
  for (auto& ex : expected_decoder_tensors(out->cfg)) {
    auto it = san.find(ex.first);
    if (it == san.end()) throw Error(Q3TTS_EFORMAT, "missing decoder tensor " + ex.first);
    if (it->second.shape != ex.second) {
      std::string got, want;
      for (auto d : it->second.shape) got += std::to_string(d) + ",";
      for (auto d : ex.second) want += std::to_string(d) + ",";
      throw Error(Q3TTS_EFORMAT, "decoder tensor " + ex.first + " has shape [" + got + "] expected [" + want + "]");
    }
    params += it->second.numel();
  }
  out->cfg.num_parameters = params;
}

// ---- encoder (row N3) ----------------------------------------------------------------------------------------------------
static void parse_encoder_config(const std::string& dir, EncoderConfig* c) {
  std::string text = read_file(dir + "/config.json");
  Json root;
  try {
    root = JsonParser(text.data(), text.size()).parse();
  } catch (const std::exception& e) {
    throw Error(Q3TTS_EFORMAT, std::string("config.json: ") + e.what());
  }
  if (root.kind != Json::Obj) throw Error(Q3TTS_EFORMAT, "config.json: top level is not an object");
  const Json* ec = root.find("encoder_config");
  if (!ec || ec->kind != Json::Obj)                                      // Qwen3.swift:433 "Speech tokenizer encoder not available"
    throw Error(Q3TTS_EFORMAT, "config.json: no encoder_config (this checkpoint has no speech-tokenizer encoder)");
  const Json& e = *ec;
  *c = EncoderConfig();
  c->valid_quantizers = geti(root, "encoder_valid_num_quantizers", 16);
  c->frame_rate = getf(e, "frame_rate", 12.5f);
  c->audio_channels = geti(e, "audio_channels", 1);
  c->codebook_dim = geti(e, "codebook_dim", 256);
  c->codebook_size = geti(e, "codebook_size", 2048);
  c->compress = geti(e, "compress", 2);
  c->dilation_growth_rate = geti(e, "dilation_growth_rate", 2);
  c->head_dim = geti(e, "head_dim", 64);
  c->hidden_size = geti(e, "hidden_size", 512);
  c->intermediate_size = geti(e, "intermediate_size", 2048);
  c->kernel_size = geti(e, "kernel_size", 7);
  c->last_kernel_size = geti(e, "last_kernel_size", 3);
  c->num_attention_heads = geti(e, "num_attention_heads", 8);
  c->num_filters = geti(e, "num_filters", 64);
  c->num_hidden_layers = geti(e, "num_hidden_layers", 8);
  c->num_key_value_heads = geti(e, "num_key_value_heads", 8);
  c->num_quantizers = geti(e, "num_quantizers", 32);
  c->num_residual_layers = geti(e, "num_residual_layers", 1);
  c->residual_kernel_size = geti(e, "residual_kernel_size", 3);
  c->rope_theta = getf(e, "rope_theta", 10000.0f);
  c->sampling_rate = geti(e, "sampling_rate", 24000);
  int32_t n = 0, r[8];
  get_list(e, "upsampling_ratios", {8, 6, 5, 4}, &n, r);
  c->n_ratios = n;
  for (int i = 0; i < n; ++i) c->ratios[i] = r[i];
  const Json* cz = e.find("use_causal_conv");
  c->use_causal_conv = (cz && cz->kind == Json::Bool) ? (cz->b ? 1 : 0) : 1;
  const Json* sc = e.find("use_conv_shortcut");
  c->use_conv_shortcut = (sc && sc->kind == Json::Bool) ? (sc->b ? 1 : 0) : 0;
  // what this implementation supports (the shipped checkpoints' values; the reference would build other graphs for the rest)
  if (!c->use_causal_conv || c->use_conv_shortcut || c->audio_channels != 1 || c->num_residual_layers != 1)
    throw Error(Q3TTS_EFORMAT, "encoder_config: only causal convs, true skip connections, mono audio and one residual layer per stage are supported");
  if (c->n_ratios < 1 || c->num_filters % 8 || c->compress != 2 || c->hidden_size % 4 || c->codebook_dim % 4 ||
      c->num_attention_heads < 1 || c->hidden_size % c->num_attention_heads || c->num_key_value_heads < 1 ||
      c->num_attention_heads % c->num_key_value_heads || c->num_quantizers < 2 || c->valid_quantizers < 2 ||
      c->valid_quantizers > c->num_quantizers || c->downsample_stride() < 1 || c->kernel_size < 1 || c->residual_kernel_size < 1)
    throw Error(Q3TTS_EFORMAT, "encoder_config: inconsistent dimensions");
}

void load_encoder_checkpoint(const std::string& dir, EncoderCheckpoint* out) {
  parse_encoder_config(dir, &out->cfg);
  const EncoderConfig& c = out->cfg;
  TensorMap raw;
  for (auto& f : list_safetensors(dir)) read_safetensors(f, &raw, "encoder.");
  static const std::pair<const char*, const char*> kSeanet[] = {   // Qwen3.swift:1517-1528 (the four-ratio layout of the shipped model)
      {"encoder.encoder.layers.0.", "encoder.encoder.init_conv1d."},
      {"encoder.encoder.layers.1.", "encoder.encoder.layers.0.residuals.0."},
      {"encoder.encoder.layers.3.", "encoder.encoder.layers.0.downsample."},
      {"encoder.encoder.layers.4.", "encoder.encoder.layers.1.residuals.0."},
      {"encoder.encoder.layers.6.", "encoder.encoder.layers.1.downsample."},
      {"encoder.encoder.layers.7.", "encoder.encoder.layers.2.residuals.0."},
      {"encoder.encoder.layers.9.", "encoder.encoder.layers.2.downsample."},
      {"encoder.encoder.layers.10.", "encoder.encoder.layers.3.residuals.0."},
      {"encoder.encoder.layers.12.", "encoder.encoder.layers.3.downsample."},
      {"encoder.encoder.layers.14.", "encoder.encoder.final_conv1d."}};
  if (c.n_ratios != 4) throw Error(Q3TTS_EFORMAT, "encoder_config.upsampling_ratios must have 4 entries (the reference's key remap names layers 0..14)");
  TensorMap& san = out->tensors;
  std::map<std::string, std::pair<const HostTensor*, const HostTensor*>> books;   // base -> (usage, sum)
  for (auto& kv : raw) {
    const std::string& key = kv.first;
    const HostTensor& value = kv.second;
    if (starts(key, "encoder.quantizer.") && has(key, ".codebook.")) {
      const size_t pos = key.find(".codebook.");
      const std::string base = key.substr(0, pos), fld = key.substr(pos + 10);
      if (fld == "embed_sum") { books[base].second = &value; continue; }
      if (fld == "cluster_usage") { books[base].first = &value; continue; }
      if (has(key, ".initialized")) continue;
    }
    std::string nk = key;
    for (auto& m : kSeanet)
      if (starts(nk, m.first)) {
        replace_all(nk, m.first, m.second);
        break;
      }
    if (has(nk, ".residuals.")) {
      replace_all(nk, ".block.1.", ".block.0.");
      replace_all(nk, ".block.3.", ".block.1.");
    }
    HostTensor nv;
    bool set = false;
    const bool is3 = value.shape.size() == 3;
    const bool seanet_conv = starts(nk, "encoder.encoder.") && !has(nk, "encoder_transformer") && !has(nk, "quantizer") &&
                             (has(nk, ".conv.weight") || has(nk, ".conv.bias"));
    if (seanet_conv) {
      replace_all(nk, ".conv.weight", ".conv.conv.weight");
      replace_all(nk, ".conv.bias", ".conv.conv.bias");
      if (nk.size() > 7 && nk.substr(nk.size() - 7) == ".weight" && is3) { nv = permute3(value, 0, 2, 1); set = true; }   // forced
    }
    if (has(nk, "encoder.encoder_transformer.layers.")) {
      replace_all(nk, "encoder.encoder_transformer.layers.", "encoder.encoder_transformer.transformer.layers.");
      replace_all(nk, ".input_layernorm.", ".norm1.");
      replace_all(nk, ".post_attention_layernorm.", ".norm2.");
      replace_all(nk, ".mlp.fc1.", ".gating.linear1.");
      replace_all(nk, ".mlp.fc2.", ".gating.linear2.");
      replace_all(nk, ".self_attn_layer_scale.", ".layer_scale_1.");
      replace_all(nk, ".mlp_layer_scale.", ".layer_scale_2.");
    }
    if (starts(nk, "encoder.downsample.conv.") && !has(nk, "encoder.downsample.conv.conv.")) {
      const bool is_w = nk.size() > 7 && nk.substr(nk.size() - 7) == ".weight";
      replace_all(nk, "encoder.downsample.conv.", "encoder.downsample.conv.conv.conv.");
      if (is_w && is3) { nv = permute3(value, 0, 2, 1); set = true; }
    }
    if (has(nk, "encoder.quantizer.")) {
      replace_all(nk, ".semantic_residual_vector_quantizer.", ".rvq_first.");
      replace_all(nk, ".acoustic_residual_vector_quantizer.", ".rvq_rest.");
      replace_all(nk, ".rvq_first.layers.", ".rvq_first.vq.layers.");
      replace_all(nk, ".rvq_rest.layers.", ".rvq_rest.vq.layers.");
    }
    const bool was_seanet_w = starts(nk, "encoder.encoder.") && !has(nk, "encoder_transformer") && !has(nk, "quantizer") &&
                              nk.size() > 17 && nk.substr(nk.size() - 17) == ".conv.conv.weight";
    const bool is_proj = (has(nk, "input_proj.weight") || has(nk, "output_proj.weight")) && has(nk, "quantizer");
    if (is_proj && is3) { nv = permute3(value, 0, 2, 1); set = true; }
    if (has(nk, "conv.weight") && is3 && !is_proj && !was_seanet_w && !is_mlx_conv_layout(value.shape)) {   // generic branch, from the ORIGINAL value
      nv = permute3(value, 0, 2, 1);
      set = true;
    }
    san[nk] = set ? std::move(nv) : value;
  }
  for (auto& b : books) {   // Qwen3.swift:1726-1748: the raw sums are kept, EncoderEuclideanCodebook divides (STE.swift:738-743)
    if (!b.second.first || !b.second.second) continue;
    std::string nb = b.first;
    replace_all(nb, ".semantic_residual_vector_quantizer.", ".rvq_first.");
    replace_all(nb, ".acoustic_residual_vector_quantizer.", ".rvq_rest.");
    replace_all(nb, ".rvq_first.layers.", ".rvq_first.vq.layers.");
    replace_all(nb, ".rvq_rest.layers.", ".rvq_rest.vq.layers.");
    san[nb + ".codebook.embeddingSum"] = *b.second.second;
    san[nb + ".codebook.clusterUsage"] = *b.second.first;
  }
  // strict validation of what encode() reads
  std::map<std::string, std::vector<int64_t>> want;
  const int nf = c.num_filters, H = c.hidden_size, hd = H / c.num_attention_heads;
  want["encoder.encoder.init_conv1d.conv.conv.weight"] = {nf, c.kernel_size, 1};
  want["encoder.encoder.init_conv1d.conv.conv.bias"] = {nf};
  int mult = 1;
  for (int li = 0; li < c.n_ratios; ++li) {
    const int ratio = c.ratios[c.n_ratios - 1 - li], dim = mult * nf, hid = dim / c.compress;
    const std::string p = "encoder.encoder.layers." + std::to_string(li);
    want[p + ".residuals.0.block.0.conv.conv.weight"] = {hid, c.residual_kernel_size, dim};
    want[p + ".residuals.0.block.0.conv.conv.bias"] = {hid};
    want[p + ".residuals.0.block.1.conv.conv.weight"] = {dim, 1, hid};
    want[p + ".residuals.0.block.1.conv.conv.bias"] = {dim};
    want[p + ".downsample.conv.conv.weight"] = {2 * dim, 2 * ratio, dim};
    want[p + ".downsample.conv.conv.bias"] = {2 * dim};
    mult *= 2;
  }
  want["encoder.encoder.final_conv1d.conv.conv.weight"] = {H, c.last_kernel_size, mult * nf};
  want["encoder.encoder.final_conv1d.conv.conv.bias"] = {H};
  for (int i = 0; i < c.num_hidden_layers; ++i) {
    const std::string p = "encoder.encoder_transformer.transformer.layers." + std::to_string(i);
    for (const char* nrm : {".norm1", ".norm2"}) { want[p + nrm + ".weight"] = {H}; want[p + nrm + ".bias"] = {H}; }
    want[p + ".self_attn.q_proj.weight"] = {H, H};
    want[p + ".self_attn.k_proj.weight"] = {c.num_key_value_heads * hd, H};
    want[p + ".self_attn.v_proj.weight"] = {c.num_key_value_heads * hd, H};
    want[p + ".self_attn.o_proj.weight"] = {H, H};
    want[p + ".gating.linear1.weight"] = {c.intermediate_size, H};
    want[p + ".gating.linear2.weight"] = {H, c.intermediate_size};
    want[p + ".layer_scale_1.scale"] = {H};
    want[p + ".layer_scale_2.scale"] = {H};
  }
  want["encoder.downsample.conv.conv.conv.weight"] = {H, 2 * c.downsample_stride(), H};
  for (int part = 0; part < 2; ++part) {
    const std::string q = std::string("encoder.quantizer.") + (part == 0 ? "rvq_first" : "rvq_rest");
    want[q + ".input_proj.weight"] = {c.codebook_dim, 1, H};
    const int n = part == 0 ? 1 : c.valid_quantizers - 1;        // the layers past encoder_valid_num_quantizers are never evaluated
    for (int i = 0; i < n; ++i) {
      want[q + ".vq.layers." + std::to_string(i) + ".codebook.embeddingSum"] = {c.codebook_size, c.codebook_dim};
      want[q + ".vq.layers." + std::to_string(i) + ".codebook.clusterUsage"] = {c.codebook_size};
    }
  }
  int64_t params = 0;
  for (auto& ex : want) {
    auto it = san.find(ex.first);
    if (it == san.end()) throw Error(Q3TTS_EFORMAT, "missing encoder tensor " + ex.first);
    if (it->second.shape != ex.second) {
      std::string got, wnt;
      for (auto d : it->second.shape) got += std::to_string(d) + ",";
      for (auto d : ex.second) wnt += std::to_string(d) + ",";
      throw Error(Q3TTS_EFORMAT, "encoder tensor " + ex.first + " has shape [" + got + "] expected [" + wnt + "]");
    }
    params += it->second.numel();
  }
  out->num_parameters = params;
}

}  // namespace q3
