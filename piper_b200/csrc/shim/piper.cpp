// Implementation of the piper.hpp-compatible surface over the C ABI (see piper.hpp in this directory).
// Behavioural contract restated from /root/reference/src/cpp/piper.cpp (line refs inline).
#include "piper.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>

#include "../../../include/piper_b200.h"

namespace piper {
namespace {

const float kMaxWavValue = 32767.0f;  // piper.cpp:29

// Decode one UTF-8 codepoint starting at s[i]; advances i.  Malformed input throws.
Phoneme decode_utf8(const std::string& s, size_t& i) {
  const unsigned char c = static_cast<unsigned char>(s[i]);
  int extra = c < 0x80 ? 0 : (c >> 5) == 0x6 ? 1 : (c >> 4) == 0xE ? 2 : (c >> 3) == 0x1E ? 3 : -1;
  if (extra < 0 || i + size_t(extra) >= s.size() + (extra == 0 ? 1 : 0)) throw std::runtime_error("invalid UTF-8");
  uint32_t cp = extra == 0 ? c : (c & (0x3F >> extra));
  for (int k = 1; k <= extra; ++k) {
    const unsigned char d = static_cast<unsigned char>(s[i + size_t(k)]);
    if ((d & 0xC0) != 0x80) throw std::runtime_error("invalid UTF-8");
    cp = (cp << 6) | (d & 0x3F);
  }
  i += size_t(extra) + 1;
  return static_cast<Phoneme>(cp);
}

size_t codepoint_count(const std::string& s) {
  size_t i = 0, n = 0;
  while (i < s.size()) {
    decode_utf8(s, i);
    ++n;
  }
  return n;
}

void check(int rc, const char* what) {
  if (rc != PB200_OK) throw std::runtime_error(std::string(what) + ": " + pb200_last_error());
}

// voice JSON -> configs (fields of piper.cpp:47-214)
void parsePhonemizeConfig(const json& root, PhonemizeConfig& pc) {
  if (const json* e = root.find("espeak"))
    if (const json* v = e->find("voice")) pc.eSpeak.voice = v->string();
  if (const json* t = root.find("phoneme_type"))
    if (t->string() == "text") pc.phonemeType = TextPhonemes;
  if (const json* m = root.find("phoneme_id_map")) {
    for (const auto& kv : m->obj) {
      if (!isSingleCodepoint(kv.first)) throw std::runtime_error("Phonemes must be one codepoint (phoneme id map)");
      const Phoneme p = getCodepoint(kv.first);
      for (const json& id : kv.second.arr) pc.phonemeIdMap[p].push_back(PhonemeId(id.number()));
    }
    // pad / bos / eos are phonemes of the voice's own map ('_', '^', '$': piper-phonemize's PhonemeIdConfig defaults, which
    // is what piper.cpp:540-560 hands to phonemes_to_ids); the struct defaults 0 / 1 / 2 only stand in when a map lacks them
    auto first_id = [&](Phoneme p, PhonemeId& out) {
      auto it = pc.phonemeIdMap.find(p);
      if (it != pc.phonemeIdMap.end() && !it->second.empty()) out = it->second.front();
    };
    first_id(U'_', pc.idPad);
    first_id(U'^', pc.idBos);
    first_id(U'$', pc.idEos);
  }
  if (const json* m = root.find("phoneme_map")) {
    if (!pc.phonemeMap) pc.phonemeMap.emplace();
    for (const auto& kv : m->obj) {
      if (!isSingleCodepoint(kv.first)) throw std::runtime_error("Phonemes must be one codepoint (phoneme map)");
      const Phoneme from = getCodepoint(kv.first);
      for (const json& to : kv.second.arr) {
        if (!isSingleCodepoint(to.string())) throw std::runtime_error("Phonemes must be one codepoint (phoneme map)");
        (*pc.phonemeMap)[from].push_back(getCodepoint(to.string()));
      }
    }
  }
}

void parseSynthesisConfig(const json& root, SynthesisConfig& sc) {
  if (const json* a = root.find("audio"))
    if (const json* r = a->find("sample_rate")) sc.sampleRate = int(r->number());
  if (const json* inf = root.find("inference")) {
    if (const json* v = inf->find("noise_scale")) sc.noiseScale = float(v->number());
    if (const json* v = inf->find("length_scale")) sc.lengthScale = float(v->number());
    if (const json* v = inf->find("noise_w")) sc.noiseW = float(v->number());
    if (const json* ps = inf->find("phoneme_silence")) {
      sc.phonemeSilenceSeconds.emplace();
      for (const auto& kv : ps->obj) {
        if (!isSingleCodepoint(kv.first)) throw std::runtime_error("Phonemes must be one codepoint (phoneme silence)");
        (*sc.phonemeSilenceSeconds)[getCodepoint(kv.first)] = float(kv.second.number());
      }
    }
  }
}

void parseModelConfig(const json& root, ModelConfig& mc) {
  mc.numSpeakers = int(root.at("num_speakers").number());
  if (const json* m = root.find("speaker_id_map")) {
    if (!mc.speakerIdMap) mc.speakerIdMap.emplace();
    for (const auto& kv : m->obj) (*mc.speakerIdMap)[kv.first] = SpeakerId(kv.second.number());
  }
}

}  // namespace

ModelSession::~ModelSession() {
  if (engine) pb200_voice_free(engine);
  engine = nullptr;
}

bool isSingleCodepoint(std::string s) { return codepoint_count(s) == 1; }

Phoneme getCodepoint(std::string s) {
  if (s.empty()) throw std::runtime_error("empty phoneme string");
  size_t i = 0;
  return decode_utf8(s, i);
}

std::string getVersion() { return pb200_version(); }

void initialize(PiperConfig& config) {
  // The reference initialises espeak-ng and libtashkeel here (piper.cpp:216-249).  Both stay on the host
  // behind config.phonemizer; nothing on the GPU side needs global initialisation.
  if (config.useTashkeel) throw std::runtime_error("libtashkeel is not available in this build");
}

void terminate(PiperConfig&) {}

void loadVoice(PiperConfig&, std::string modelPath, std::string modelConfigPath, Voice& voice,
               std::optional<SpeakerId>& speakerId, bool useCuda) {
  std::ifstream f(modelConfigPath);
  if (!f) throw std::runtime_error("cannot open voice config '" + modelConfigPath + "'");
  std::stringstream ss;
  ss << f.rdbuf();
  voice.configRoot = minijson::parse(ss.str());
  parsePhonemizeConfig(voice.configRoot, voice.phonemizeConfig);
  parseSynthesisConfig(voice.configRoot, voice.synthesisConfig);
  parseModelConfig(voice.configRoot, voice.modelConfig);
  if (voice.modelConfig.numSpeakers > 1) voice.synthesisConfig.speakerId = speakerId ? speakerId : SpeakerId(0);
  if (!useCuda)
    throw std::runtime_error("piper_b200 has no CPU execution path: pass --cuda / useCuda = true");
  check(pb200_voice_load(modelPath.c_str(), voice.session.device, &voice.session.engine), "loadVoice");
}

void synthesize(std::vector<PhonemeId>& phonemeIds, SynthesisConfig& sc, ModelSession& session,
                std::vector<int16_t>& audioBuffer, SynthesisResult& result) {
  if (!session.engine) throw std::runtime_error("voice is not loaded");
  const float scales[3] = {sc.noiseScale, sc.lengthScale, sc.noiseW};   // piper.cpp:345-347
  int64_t sid = sc.speakerId.value_or(0);
  pb200_noise noise{nullptr, nullptr, 0, session.noiseSeed};
  session.noiseSeed = session.noiseSeed * 6364136223846793005ull + 1442695040888963407ull;
  const float* audio = nullptr;
  int64_t audioCount = 0;
  const auto t0 = std::chrono::steady_clock::now();               // timed like piper.cpp:385-395
  check(pb200_synthesize(session.engine, phonemeIds.data(), int64_t(phonemeIds.size()), scales,
                         sc.speakerId ? &sid : nullptr, &noise, &audio, &audioCount, nullptr),
        "synthesize");
  result.inferSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  result.audioSeconds = double(audioCount) / double(sc.sampleRate);
  result.realTimeFactor = result.audioSeconds > 0 ? result.inferSeconds / result.audioSeconds : 0.0;

  // peak-normalise to int16 exactly as piper.cpp:411-431
  float peak = 0.01f;
  for (int64_t i = 0; i < audioCount; ++i) peak = std::max(peak, std::fabs(audio[i]));
  const float scale = kMaxWavValue / std::max(0.01f, peak);
  audioBuffer.reserve(audioBuffer.size() + size_t(audioCount));
  const float lo = float(std::numeric_limits<int16_t>::min()), hi = float(std::numeric_limits<int16_t>::max());
  for (int64_t i = 0; i < audioCount; ++i)
    audioBuffer.push_back(static_cast<int16_t>(std::clamp(audio[i] * scale, lo, hi)));
  pb200_release(session.engine, audio);
}

void phonemes_to_ids(const std::vector<Phoneme>& phonemes, const PhonemizeConfig& config,
                     std::vector<PhonemeId>& ids, std::map<Phoneme, std::size_t>& missing) {
  ids.push_back(config.idBos);
  if (config.interspersePad) ids.push_back(config.idPad);
  for (Phoneme p : phonemes) {
    auto it = config.phonemeIdMap.find(p);
    if (it == config.phonemeIdMap.end()) {
      ++missing[p];
      continue;
    }
    ids.insert(ids.end(), it->second.begin(), it->second.end());
    if (config.interspersePad) ids.push_back(config.idPad);
  }
  ids.push_back(config.idEos);
}

void phonemesToAudio(PiperConfig&, Voice& voice, std::vector<std::vector<Phoneme>>& sentences,
                     std::vector<int16_t>& audioBuffer, SynthesisResult& result,
                     const std::function<void()>& audioCallback) {
  SynthesisConfig& sc = voice.synthesisConfig;
  const size_t sentenceSilence =
      sc.sentenceSilenceSeconds > 0 ? size_t(sc.sentenceSilenceSeconds * sc.sampleRate * sc.channels) : 0;
  std::map<Phoneme, std::size_t> missing;
  std::vector<PhonemeId> ids;
  for (std::vector<Phoneme>& sentence : sentences) {
    // optional split into phrases at phonemes that carry extra silence (piper.cpp:508-537)
    std::vector<std::vector<Phoneme>> phrases(1);
    std::vector<size_t> silence;
    if (sc.phonemeSilenceSeconds) {
      for (Phoneme p : sentence) {
        phrases.back().push_back(p);
        auto it = sc.phonemeSilenceSeconds->find(p);
        if (it != sc.phonemeSilenceSeconds->end()) {
          silence.push_back(size_t(it->second * sc.sampleRate * sc.channels));
          phrases.emplace_back();
        }
      }
    } else {
      phrases.back() = sentence;
    }
    silence.resize(phrases.size(), 0);
    for (size_t p = 0; p < phrases.size(); ++p) {
      if (phrases[p].empty()) continue;
      ids.clear();
      phonemes_to_ids(phrases[p], voice.phonemizeConfig, ids, missing);
      SynthesisResult r;
      synthesize(ids, sc, voice.session, audioBuffer, r);
      audioBuffer.insert(audioBuffer.end(), silence[p], int16_t(0));
      result.audioSeconds += r.audioSeconds;
      result.inferSeconds += r.inferSeconds;
    }
    audioBuffer.insert(audioBuffer.end(), sentenceSilence, int16_t(0));
    if (audioCallback) {  // per-sentence streaming hook; the callback must copy (piper.cpp:591-595)
      audioCallback();
      audioBuffer.clear();
    }
  }
  if (result.audioSeconds > 0) result.realTimeFactor = result.inferSeconds / result.audioSeconds;
}

void textToAudio(PiperConfig& config, Voice& voice, std::string text, std::vector<int16_t>& audioBuffer,
                 SynthesisResult& result, const std::function<void()>& audioCallback) {
  std::vector<std::vector<Phoneme>> sentences;
  if (voice.phonemizeConfig.phonemeType == TextPhonemes) {
    // phonemize_codepoints: one "sentence" of codepoints.  piper-phonemize (a dependency that is not in the reference tree;
    // CodepointsPhonemeConfig, casing = CASING_FOLD by default) case-folds and NFD-normalises the text first.  Folding is
    // done here for ASCII and the Latin-1 letters (one-to-one lower-casing); full Unicode case folding and decomposition
    // need the Unicode tables and stay with the host: pass text that is already case-folded / NFD if the voice needs it.
    sentences.emplace_back();
    for (size_t i = 0; i < text.size();) {
      Phoneme c = decode_utf8(text, i);
      if ((c >= U'A' && c <= U'Z') || (c >= 0xC0 && c <= 0xDE && c != 0xD7)) c += 32;
      sentences.back().push_back(c);
    }
  } else {
    if (!config.phonemizer)
      throw std::runtime_error(
          "espeak-ng phonemization is a host hook (PiperConfig::phonemizer) and none is installed; "
          "use phonemesToAudio / phoneme ids for pre-phonemized input");
    config.phonemizer(text, voice.phonemizeConfig.eSpeak, sentences);
  }
  phonemesToAudio(config, voice, sentences, audioBuffer, result, audioCallback);
}

void writeWavHeader(int sampleRate, int sampleWidth, int channels, uint32_t numSamples, std::ostream& out) {
  // 44-byte canonical PCM header (src/cpp/wavfile.hpp:6-38)
  auto u32 = [&](uint32_t v) { out.write(reinterpret_cast<const char*>(&v), 4); };
  auto u16 = [&](uint16_t v) { out.write(reinterpret_cast<const char*>(&v), 2); };
  const uint32_t dataSize = numSamples * uint32_t(sampleWidth) * uint32_t(channels);
  out.write("RIFF", 4); u32(dataSize + 36);
  out.write("WAVE", 4); out.write("fmt ", 4); u32(16); u16(1); u16(uint16_t(channels)); u32(uint32_t(sampleRate));
  u32(uint32_t(sampleRate * sampleWidth * channels)); u16(uint16_t(sampleWidth * channels)); u16(16);
  out.write("data", 4); u32(dataSize);
}

void textToWavFile(PiperConfig& config, Voice& voice, std::string text, std::ostream& audioFile,
                   SynthesisResult& result) {
  std::vector<int16_t> audio;
  textToAudio(config, voice, text, audio, result, nullptr);
  const SynthesisConfig& sc = voice.synthesisConfig;
  writeWavHeader(sc.sampleRate, sc.sampleWidth, sc.channels, uint32_t(audio.size()), audioFile);
  audioFile.write(reinterpret_cast<const char*>(audio.data()), std::streamsize(sizeof(int16_t) * audio.size()));
}

void phonemeIdsToWavFile(Voice& voice, std::vector<PhonemeId>& phonemeIds, std::ostream& audioFile,
                         SynthesisResult& result) {
  std::vector<int16_t> audio;
  synthesize(phonemeIds, voice.synthesisConfig, voice.session, audio, result);
  const SynthesisConfig& sc = voice.synthesisConfig;
  writeWavHeader(sc.sampleRate, sc.sampleWidth, sc.channels, uint32_t(audio.size()), audioFile);
  audioFile.write(reinterpret_cast<const char*>(audio.data()), std::streamsize(sizeof(int16_t) * audio.size()));
}

}  // namespace piper
