{{- define "infomesh.name" -}}{{ .Release.Name | trunc 50 | trimSuffix "-" }}{{- end -}}
{{- define "infomesh.labels" -}}
app.kubernetes.io/name: infomesh
app.kubernetes.io/instance: {{ .Release.Name }}
app.kubernetes.io/version: {{ .Chart.AppVersion | quote }}
{{- end -}}
{{- define "infomesh.secret" -}}{{ default (printf "%s-api" (include "infomesh.name" .)) .Values.apiKeySecret }}{{- end -}}
{{- define "infomesh.image" -}}{{ .Values.image.repository }}:{{ .Values.image.tag }}{{- end -}}
